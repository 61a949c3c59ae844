"""ADI+mSDI (IFS, 4-d cubes) full-frame PCA on the B200: the ``scale_list`` branches of ``pca``.

Reference: ``src/vip_hci/psfsub/pca_fullfr.py`` -- ``_adimsdi_doublepca`` :1245-1475,
``_adimsdi_doublepca_ifs`` :1478-1549, ``_adimsdi_singlepca`` :1038-1242; rescaling through
``cube_rescaling_wavelengths`` (``preproc/rescaling.py:324-475``).

Double pass: per ADI frame, the z spectral channels are rescaled by ``scale_list`` (speckles aligned,
planets move radially), a PCA across channels removes ``ncomp[0]`` components, the residuals are
descaled, collapsed over channels and cropped; the resulting (n, H, W) cube then goes through the
ordinary ADI PCA (``ncomp[1]``), derotation and collapse.  Frames are independent in stage 1 and are
processed in chunks on the device; the (de)scaling is the separable operator of
``preproc/rescaling.py`` applied with batched GEMMs.
"""
import numpy as np
import torch

from .. import kernels
from .._device import require_cuda, to_device_f32
from ..preproc.derotation import derotate_device
from ..preproc.parangles import check_pa_vector
from ..preproc.rescaling import crop_window, padded_size, rescale_operator
from ..preproc.subsampling import collapse_device
from ..var.coords import frame_center
from ..var.shapes import circle_mask
from .pca_fullfr import project_subtract_device

# device memory budget (bytes) for one chunk of rescaled multispectral frames
_CHUNK_BYTES = 2 << 30


class RescaleOps:
    """Forward / inverse rescaling operators of every channel, on the device.

    forward:  (z, S', S')  real and imaginary parts stacked as W = [Lr; Li] (z, 2S', S')
    inverse:  rows restricted to the final crop window -> (z, 2H, S')

    Instances are cached per (scale_list, frame size, device): building and uploading the 2 z operators (and, on
    first use, their bf16x3 planes for the tensor-core products) costs more than a whole config-4 frame chunk."""

    _cache = {}

    def __new__(cls, scale_list, size, device):
        key = (np.asarray(scale_list, dtype=np.float64).tobytes(), int(size), str(device))
        hit = cls._cache.get(key)
        if hit is not None:
            return hit
        self = super().__new__(cls)
        self._build(scale_list, size, device)
        if len(cls._cache) >= 4:
            cls._cache.pop(next(iter(cls._cache)))
        cls._cache[key] = self
        return self

    def __init__(self, scale_list, size, device):
        pass

    def _build(self, scale_list, size, device):
        scale_list = np.asarray(scale_list, dtype=np.float64)
        self.z = scale_list.shape[0]
        self.size = size
        max_sc = float(np.amax(scale_list))
        self.big = padded_size(size, max_sc)
        self.pad = (self.big - size) // 2
        fwd = [rescale_operator(self.big, float(s)) for s in scale_list]
        inv = [rescale_operator(self.big, float(1.0 / s)) for s in scale_list]
        # inverse: collapse and crop commute (per-pixel statistics), so the crop is folded in the operator
        if max_sc > 1 and self.big > size:
            c = frame_center((self.big, self.big))[0]
            y0, y1 = crop_window(self.big, size, c)
        else:
            y0, y1 = 0, self.big
        self.out = y1 - y0
        inv = [L[y0:y1] for L in inv]

        def stack(ops):
            W = np.stack([np.concatenate((L.real, L.imag), axis=0) for L in ops]).astype(np.float32)
            return torch.from_numpy(W).to(device)
        self.Wf = stack(fwd)          # (z, 2*big, big)
        self.Wi = stack(inv)          # (z, 2*out, big)

    @staticmethod
    def apply(X, W, z):
        """X (B, S, S) fp32 with B a multiple of z (channel = b % z); W (z, 2*So, S).
        Returns Re(L X L^T) (B, So, So) with L = W[:So] + i W[So:].

        Default: two products on the tensor cores (``kernels.gemm_tc``, tcgen05 with fp32-grade bf16x3 operands):
        U = [Lr; Li] X^T written by the first kernel's epilogue directly as the bf16x3 planes of
        [Lr X^T | Li X^T] (So x 2S per frame), then  Y = [Lr | -Li] U^T  with K = 2S.  The CUDA-core fp32 GEMM
        (``kernels.gemm``) remains for odd sizes and ``VIP_B200_RESCALE_TC=0``."""
        B, S, _ = X.shape
        So = W.shape[1] // 2
        if _rescale_tc() and S % 2 == 0 and So % 2 == 0 and B % z == 0 and B <= 65535:
            ops = getattr(W, "_vb_tc_planes", None)
            if ops is None:
                A1 = kernels.split3(W.reshape(z * 2 * So, S).contiguous())
                A2 = kernels.split3(torch.cat((W[:, :So], -W[:, So:]), dim=2).reshape(z * So, 2 * S).contiguous())
                ops = W._vb_tc_planes = (A1, A2)
            A1, A2 = ops
            Xp = kernels.split3(X.reshape(B * S, S))
            U = kernels.planes3_empty(B * So, 2 * S, X.device)
            kernels.gemm_tc(A1, z, Xp, 2 * So, S, B, out_planes=U, msplit=So)
            Y = torch.empty((B, So, So), dtype=torch.float32, device=X.device)
            return kernels.gemm_tc(A2, z, U, So, So, B, out=Y)
        U = torch.empty((B, S, 2 * So), dtype=torch.float32, device=X.device)
        kernels.gemm(X, W, U, trans_b=True, b_mod=z)                       # U = X [Lr; Li]^T
        Y = torch.empty((B, So, So), dtype=torch.float32, device=X.device)
        kernels.gemm(W[:, :So], U[:, :, :So], Y, a_mod=z)                  # Lr (X Lr^T)
        kernels.gemm(W[:, So:], U[:, :, So:], Y, a_mod=z, alpha=-1.0, beta=1.0)   # - Li (X Li^T)
        return Y


def _rescale_tc():
    import os
    return os.environ.get("VIP_B200_RESCALE_TC", "1") != "0"


def _batched_channel_pca(resc, ncomp, mask_center_px):
    """PCA across the z channels of every frame of a chunk, several frames per launch (the default configuration of
    ``_adimsdi_doublepca_ifs``: integer ``ncomp``, no ``scaling``, exact SVD): with G_f = M_f M_f^T (z x z) and E_k its
    k leading eigenvectors, the residuals of ``_project_subtract`` are R_f = (I - E_k E_k^T) M_f.  The channel
    matrices of a group of frames are stacked ((g z) x p, g z <= 256): ONE Gramian on the tensor cores serves the
    group (its diagonal z x z blocks; exact bf16x3 products, fp64 accumulation -- an fp32 Gramian would scramble the
    noise-level eigenvectors), ONE launch of the batched sub-Gramian eigensolver of the annular path solves the g z
    problems (problem (f, c) returns column c of E_k E_k^T = E diag(1/lambda) E^T G e_c) and one product applies
    them -- the same scheme as the spectral pass of the annular ADI+mSDI (``annular.py``).  At BASELINE config 4 the
    frame-by-frame route spent more time in launch latencies and Python than in its kernels
    (profiles/r04_summary.md).  resc (F, z, S, S) -> residuals (F, z, S, S)."""
    F, z, S, _ = resc.shape
    p = S * S
    M = resc.reshape(F * z, p)
    if mask_center_px:
        mask = torch.as_tensor(circle_mask((S, S), mask_center_px).reshape(-1)).to(M.device)
        M = M.masked_fill(mask[None, :], 0.0)
    k = int(ncomp)
    if k > min(z, p):
        msg = "{} PCs cannot be obtained from a matrix with size [{},{}]."
        msg += " Increase the size of the patches or request less PCs"
        raise RuntimeError(msg.format(k, z, p))
    group = max(1, 256 // z)
    R = M.clone()
    ar = torch.arange(z, dtype=torch.int32, device=M.device)
    for g0 in range(0, F, group):
        g1 = min(F, g0 + group)
        Fg = g1 - g0
        Ag = M[g0 * z:g1 * z]
        G = kernels.gram(Ag)
        base = torch.arange(Fg, device=M.device, dtype=torch.int32).repeat_interleave(z) * z
        idx = (base[:, None] + ar[None, :]).contiguous()                        # library of (f, c): block f
        lens = torch.full((Fg * z,), z, dtype=torch.int32, device=M.device)
        # the direct (tridiagonalisation) solver for every problem: 39 x 39 matrices cost it ~50 us each, it cannot
        # fail to converge, and -- unlike the subspace iteration with its fallback list -- needs no host round trip,
        # so the host keeps enqueueing (and uploading the next chunk) while the GPU works
        Wt, _ = kernels.annular_weights(G, idx, lens, torch.arange(Fg * z, dtype=torch.int32, device=M.device), k,
                                        force_direct=True)
        # W is block diagonal (the library of (f, c) is frame f): apply the Fg diagonal z x z blocks as a batch
        P = Wt.reshape(Fg, z, Fg, z).diagonal(dim1=0, dim2=2).permute(2, 0, 1).contiguous()      # (Fg, z, z)
        kernels.gemm(P, Ag.reshape(Fg, z, p), R[g0 * z:g1 * z].reshape(Fg, z, p), alpha=-1.0, beta=1.0)   # R = M - P M
    return R.reshape(F, z, S, S)


def _stage1_frames(cube_dev, frames, ops, ncomp_ifs, scaling, mask_center_px, svd_mode, collapse_ifs,
                   ifs_range):
    """``_adimsdi_doublepca_ifs`` for a chunk of ADI frames: (z, n, H, W) device cube -> (F, H, W)."""
    from .pca_fullfr import _EXACT_MODES, _mode_name
    z, n, H, W = cube_dev.shape
    F = len(frames)
    i0, i1 = ifs_range
    ms = cube_dev[:, frames].permute(1, 0, 2, 3).contiguous()              # (F, z, H, W)
    if ncomp_ifs is None:
        return torch.stack([collapse_device(ms[f, i0:i1], "median") for f in range(F)])
    if ops.pad:
        ms = torch.nn.functional.pad(ms, (ops.pad,) * 4, mode="reflect")   # np.pad(..., 'reflect')
    S = ops.big
    resc = RescaleOps.apply(ms.reshape(F * z, S, S), ops.Wf, z).reshape(F, z, S, S)
    if (isinstance(ncomp_ifs, (int, np.integer)) and scaling is None and _mode_name(svd_mode) in _EXACT_MODES
            and z <= 256 and ncomp_ifs <= 24 and _batched_stage1()):
        res = _batched_channel_pca(resc, ncomp_ifs, mask_center_px)
    else:
        res = torch.empty_like(resc)
        for f in range(F):                                                 # PCA across the z channels
            res[f] = project_subtract_device(resc[f], ncomp_ifs, scaling, mask_center_px, svd_mode)
    desc = RescaleOps.apply(res.reshape(F * z, S, S), ops.Wi, z).reshape(F, z, ops.out, ops.out)
    # collapse over the selected channels of every frame in one launch: (zr, F * out * out) along axis 0
    sel = desc[:, i0:i1].permute(1, 0, 2, 3).reshape(i1 - i0, F * ops.out, ops.out)
    out = collapse_device(sel.contiguous(), collapse_ifs).reshape(F, ops.out, ops.out)
    if out.dtype != torch.float32:
        out = out.float()
    if mask_center_px:
        mask = torch.as_tensor(circle_mask((ops.out, ops.out), mask_center_px)).to(out.device)
        out = out.masked_fill(mask[None], 0.0)
    return out


class _FrameFeeder:
    """Frames [f0, f1) of every channel of a (z, n, H, W) host cube (a reference cube, if any, continues the frame
    axis) -> (z, F, H, W) device tensors, uploaded on a copy stream ``depth`` chunks ahead of the chunk whose kernels
    are being enqueued (``vb_memcpy2d_h2d_staged``: strided rows gathered into pinned buffers by worker threads).  At
    BASELINE config 4 the 3.07 GB upload (110 ms from a pageable numpy array) disappears behind the first pass."""

    def __init__(self, cube, cube_ref, chunks, dev, depth=3):
        self.cubes = [cube] + ([cube_ref] if cube_ref is not None else [])
        self.n = cube.shape[1]
        self.chunks, self.dev, self.depth = chunks, dev, depth
        self.pending = {}
        self.issued = 0
        self.stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None

    def _upload_one(self, cube, f0, f1):
        z, n, H, W = cube.shape
        fast = (self.stream is not None and isinstance(cube, np.ndarray) and cube.dtype == np.float32
                and cube.flags["C_CONTIGUOUS"])
        if not fast:
            part = cube[:, f0:f1]
            return to_device_f32(np.ascontiguousarray(part) if isinstance(part, np.ndarray) else part, self.dev)
        from .. import _cabi
        from .._device import stream_ptr
        out = torch.empty((z, f1 - f0, H, W), dtype=torch.float32, device=self.dev)
        frame_bytes = H * W * 4
        _cabi.check(_cabi.lib().vb_memcpy2d_h2d_staged(int(out.data_ptr()), int(cube.ctypes.data) + f0 * frame_bytes,
                                                       n * frame_bytes, (f1 - f0) * frame_bytes, z, stream_ptr()),
                    "vb_memcpy2d_h2d_staged")
        return out

    def _upload(self, i):
        f0, f1 = self.chunks[i]
        parts = []
        if f0 < self.n:
            parts.append(self._upload_one(self.cubes[0], f0, min(f1, self.n)))
        if f1 > self.n:
            parts.append(self._upload_one(self.cubes[1], max(f0, self.n) - self.n, f1 - self.n))
        return parts[0] if len(parts) == 1 else torch.cat(parts, dim=1)

    def get(self, i):
        """Device tensor of chunk i (ready on the current stream); keeps the uploads ``depth`` chunks ahead."""
        while self.issued < min(len(self.chunks), i + self.depth):
            j = self.issued
            if self.stream is None:
                self.pending[j] = (self._upload(j), None)
            else:
                with torch.cuda.stream(self.stream):
                    t = self._upload(j)
                    ev = torch.cuda.Event()
                    ev.record(self.stream)
                self.pending[j] = (t, ev)
            self.issued += 1
        t, ev = self.pending.pop(i)
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)
            t.record_stream(torch.cuda.current_stream())
        return t


def _batched_stage1():
    import os
    return os.environ.get("VIP_B200_SDI_BATCHED", "1") != "0"


def adimsdi_doublepca_device(cube, angle_list, scale_list, ncomp, scaling=None, mask_center_px=None,
                             svd_mode="lapack", collapse="median", collapse_ifs="mean",
                             ifs_collapse_range="all", weights=None, verbose=False, cube_ref=None,
                             ref_strategy="RSDI", source_xy=None, delta_rot=None, fwhm=4, min_frames_pca=10,
                             max_frames_pca=None, cube_sig=None, **rot_options):
    """``_adimsdi_doublepca`` (``pca_fullfr.py:1245-1475``) on the device.

    ``cube_ref`` (z, nr, H, W): its frames go through the same spectral first pass (:1278-1282, :1326) and form the
    reference library of the second, rotational pass (``ref_strategy`` 'RSDI', :1392-1404).  ``source_xy``: second
    pass frame by frame with a PA-rejection library (:1407-1460).  ``cube_sig`` (n, H, W): subtracted from the
    first-pass frames before the second decomposition.
    Returns (res_cube_channels (n+nr,H,W), residuals_cube_channels_ (n,H,W), frame (H,W)) as CUDA tensors."""
    z, n, y_in, x_in = cube.shape
    nr = 0
    if cube_ref is not None:
        if np.ndim(cube_ref) != 4:
            raise TypeError("Ref cube has wrong format for 4d input cube")
        nr = cube_ref.shape[1]
    if not isinstance(ncomp, tuple):
        raise TypeError("`ncomp` must be a tuple when a double pass PCA is performed")
    ncomp_ifs, ncomp_adi = ncomp
    angle_list = check_pa_vector(np.asarray(angle_list))
    if angle_list.shape[0] != n:
        raise ValueError("Angle list vector has wrong length. It must equal the number frames in the cube")
    if scale_list is None:
        raise ValueError("Scaling factors vector must be provided")
    scale_list = np.asarray(scale_list)
    if scale_list.ndim > 1:
        raise ValueError("Scaling factors vector is not 1d")
    if scale_list.shape[0] != z:
        raise ValueError("Scaling factors vector has wrong length")
    if y_in != x_in:
        raise ValueError("FFT scaling only supports square input arrays")
    if not isinstance(scaling, tuple):
        scaling = (scaling, scaling)
    if ncomp_ifs is not None and ncomp_ifs > z:
        ncomp_ifs = min(ncomp_ifs, z)
        print("Number of PCs too high (max PCs={}), using {} PCs instead".format(z, ncomp_ifs))
    ifs_range = (0, z) if ifs_collapse_range == "all" else (int(ifs_collapse_range[0]), int(ifs_collapse_range[1]))

    dev = require_cuda()
    ops = RescaleOps(scale_list, y_in, dev)
    per_frame = 4 * z * ops.big * ops.big * 4 * 2          # rescaled + U + residuals + descaled (upper bound)
    chunk = max(1, min(n + nr, int(_CHUNK_BYTES // per_frame)))
    chunks = [(f0, min(n + nr, f0 + chunk)) for f0 in range(0, n + nr, chunk)]
    feeder = _FrameFeeder(cube, cube_ref if nr else None, chunks, dev)      # the reference frames follow the cube's
    parts = []
    for i, (f0, f1) in enumerate(chunks):
        sub = feeder.get(i)                                                  # (z, F, H, W), uploaded ahead
        parts.append(_stage1_frames(sub, list(range(f1 - f0)), ops, ncomp_ifs, scaling[0], mask_center_px, svd_mode,
                                    collapse_ifs, ifs_range))
    res_channels = torch.cat(parts)                          # (n, H, W)
    if verbose:
        print("First PCA stage exploiting spectral variability done; {} ADI frames".format(n))

    mask_val = float(rot_options.get("mask_val", np.nan))
    interp_zeros = bool(rot_options.get("interp_zeros", False))
    sci = res_channels[:n]
    ref = res_channels[n:] if nr else None
    sig_dev = to_device_f32(cube_sig, dev) if cube_sig is not None else None
    if ncomp_adi is None:
        res2 = sci
    else:
        if ncomp_adi > n + nr:
            ncomp_adi = n + nr
            print("Number of PCs too high, using  maximum of {} PCs instead".format(n))
        if source_xy is None and nr and "A" in str(ref_strategy):
            # upstream decomposes the n + nr first-pass frames and then derotates all of them with the n angles
            # (pca_fullfr.py:1379-1391, 1462: IndexError in cube_derotate): nothing to reproduce
            raise IndexError("index {} is out of bounds for axis 0 with size {} (ref_strategy='ARSDI' with "
                             "adimsdi='double' fails the same way in the reference)".format(n, n))
        if source_xy is None:
            res2 = project_subtract_device(sci, ncomp_adi, scaling[1], mask_center_px, svd_mode,
                                           cube_ref_dev=ref, cube_sig_dev=sig_dev, in_dtype=np.float64)
        else:
            from .pca_fullfr import pa_rejection_residuals_device
            res2, _, _ = pa_rejection_residuals_device(sci, ref, sig_dev, angle_list, ncomp_adi, scaling[1],
                                                       mask_center_px, svd_mode, source_xy, delta_rot, fwhm,
                                                       min_frames_pca, max_frames_pca, in_dtype=np.float64)
    res_der = derotate_device(res2, -angle_list, mask_val=mask_val, interp_zeros=interp_zeros)
    frame = collapse_device(res_der, mode=collapse, w=weights)
    return res_channels, res_der, frame


class _SinglePass:
    """Shared pieces of the single-pass ADI+mSDI PCA: the rescaled (and cropped) stack of all channels of all ADI
    frames (``pca_fullfr.py:1083-1097``) and its inverse -- descaling of the selected channels of every ADI frame,
    collapse over them, crop to the input size (``scwave(..., inverse=True)``, :1161-1175 / ``utils_pca.py:203-221``)."""

    def __init__(self, z, n, y_in, scale_list, ifs_collapse_range, crop_ifs, dev):
        self.z, self.n, self.y_in, self.dev = z, n, y_in, dev
        self.scale_list = scale_list
        self.i0, self.i1 = ((0, z) if ifs_collapse_range == "all"
                            else (int(ifs_collapse_range[0]), int(ifs_collapse_range[1])))
        self.ops = RescaleOps(scale_list, y_in, dev)
        S = self.ops.big
        # crop_ifs: cube_crop_frames(cube_resc, size=y_in) about the centre of the padded frame
        if crop_ifs and S > y_in:
            c = frame_center((S, S))[0]
            self.c0, self.c1 = crop_window(S, y_in, c)
        else:
            self.c0, self.c1 = 0, S
        self.Sc = self.c1 - self.c0
        per_frame = 3 * z * S * S * 4 * 2
        self.chunk = max(1, int(_CHUNK_BYTES // per_frame))
        self._Wi = None

    def rescaled_stack(self, cube_dev):
        """(z, m, H, W) device cube -> (m*z, Sc, Sc), frame index = i*z + channel."""
        z, m = cube_dev.shape[:2]
        S, c0, c1 = self.ops.big, self.c0, self.c1
        big = torch.empty((m * z, self.Sc, self.Sc), dtype=torch.float32, device=self.dev)
        for f0 in range(0, m, self.chunk):
            f1 = min(m, f0 + self.chunk)
            ms = cube_dev[:, f0:f1].permute(1, 0, 2, 3).contiguous()                # (F, z, H, W)
            if self.ops.pad:
                ms = torch.nn.functional.pad(ms, (self.ops.pad,) * 4, mode="reflect")
            resc = RescaleOps.apply(ms.reshape((f1 - f0) * z, S, S), self.ops.Wf, z)
            big[f0 * z:f1 * z] = resc[:, c0:c1, c0:c1]
        return big

    def inverse_ops(self):
        # inverse rescaling acts on the (possibly cropped) frames: operators of size Sc, crop to the input size
        if self._Wi is None:
            Sc, y_in = self.Sc, self.y_in
            inv = [rescale_operator(Sc, float(1.0 / s)) for s in self.scale_list[self.i0:self.i1]]
            if Sc > y_in:
                c = frame_center((Sc, Sc))[0]
                r0, r1 = crop_window(Sc, y_in, c)
                inv = [L[r0:r1] for L in inv]
            W = np.stack([np.concatenate((L.real, L.imag), axis=0) for L in inv]).astype(np.float32)
            self._Wi = torch.from_numpy(W).to(self.dev)
        return self._Wi

    def descale_collapse(self, res_cube, collapse_mode, want_desc=False):
        """Residuals of the stack (n*z, Sc, Sc) -> (cube_desc_residuals (zr, n, H, W) or None, (n, H, W))."""
        z, n, y_in, Sc, i0, i1 = self.z, self.n, self.y_in, self.Sc, self.i0, self.i1
        Wi = self.inverse_ops()
        zr = i1 - i0
        desc = torch.empty((zr, n, y_in, y_in), dtype=torch.float32, device=self.dev) if want_desc else None
        resadi = torch.empty((n, y_in, y_in), dtype=torch.float32, device=self.dev)
        for f0 in range(0, n, self.chunk):
            f1 = min(n, f0 + self.chunk)
            F = f1 - f0
            sel = res_cube[f0 * z:f1 * z].reshape(F, z, Sc, Sc)[:, i0:i1].reshape(F * zr, Sc, Sc).contiguous()
            d = RescaleOps.apply(sel, Wi, zr).reshape(F, zr, y_in, y_in)
            if want_desc:
                desc[:, f0:f1] = d.permute(1, 0, 2, 3)
            for f in range(F):
                resadi[f0 + f] = collapse_device(d[f], collapse_mode).float()
        return desc, resadi


def _check_single_inputs(cube, angle_list, scale_list):
    z, n, y_in, x_in = cube.shape
    angle_list = check_pa_vector(np.asarray(angle_list))
    if angle_list.shape[0] != n:
        raise ValueError("Angle list vector has wrong length. It must equal the number frames in the cube")
    if scale_list is None:
        raise ValueError("`scale_list` must be provided")
    scale_list = np.asarray(scale_list)
    if scale_list.ndim != 1:
        raise TypeError("Input array (scale_list) is not a 1d array")
    if scale_list.shape[0] != z:
        raise ValueError("`scale_list` has wrong length")
    if y_in != x_in:
        raise ValueError("FFT scaling only supports square input arrays")
    return angle_list, scale_list


def adimsdi_singlepca_device(cube, angle_list, scale_list, ncomp, scaling=None, mask_center_px=None,
                             svd_mode="lapack", collapse="median", collapse_ifs="mean",
                             ifs_collapse_range="all", crop_ifs=True, weights=None, verbose=False, cube_ref=None,
                             **rot_options):
    """``_adimsdi_singlepca`` (``pca_fullfr.py:1038-1242``) for a scalar ``ncomp`` on the device: every
    channel of every ADI frame is rescaled (speckles aligned), ONE PCA runs over the z*n rescaled frames,
    the residuals are descaled and collapsed over the channels of each ADI frame, then derotated and
    collapsed over time.  ``cube_ref`` (z, nr, H, W): rescaled the same way, its z*nr frames are the PCA library
    (:1099-1119; for ``ref_strategy='ARSDI'`` the caller passes cube and reference concatenated, :504-508).
    Returns (cube_allfr_residuals (z*n,S,S), cube_desc_residuals (zr,n,H,W),
    cube_adi_residuals (n,H,W), frame (H,W)) as CUDA tensors (frame index of the big cube = i*z + channel)."""
    z, n, y_in, x_in = cube.shape
    angle_list, scale_list = _check_single_inputs(cube, angle_list, scale_list)
    if not np.isscalar(ncomp):
        raise TypeError("`ncomp` must be an int, float, tuple or list for single-pass PCA")
    dev = require_cuda()
    sp = _SinglePass(z, n, y_in, scale_list, ifs_collapse_range, crop_ifs, dev)
    big = sp.rescaled_stack(to_device_f32(cube, dev))
    big_ref = sp.rescaled_stack(to_device_f32(cube_ref, dev)) if cube_ref is not None else None
    if verbose:
        print("{} total frames".format(n * z))
        print("Performing single-pass PCA")
    res_cube = project_subtract_device(big, ncomp, scaling, mask_center_px, svd_mode, cube_ref_dev=big_ref,
                                       in_dtype=np.float64)
    desc, resadi = sp.descale_collapse(res_cube, collapse_ifs, want_desc=True)
    mask_val = float(rot_options.get("mask_val", np.nan))
    interp_zeros = bool(rot_options.get("interp_zeros", False))
    der = derotate_device(resadi, -angle_list, mask_val=mask_val, interp_zeros=interp_zeros)
    if mask_center_px:
        mask = torch.as_tensor(circle_mask((y_in, x_in), mask_center_px)).to(dev)
        der = der.masked_fill(mask[None], 0.0)
    frame = collapse_device(der, mode=collapse, w=weights)
    return res_cube, desc, resadi, frame


def adimsdi_singlepca_grid_device(cube, angle_list, scale_list, range_pcs, scaling=None, mask_center_px=None,
                                  svd_mode="lapack", collapse="median", ifs_collapse_range="all", crop_ifs=True,
                                  weights=None, verbose=False, **rot_options):
    """``_adimsdi_singlepca`` with a tuple / list ``ncomp`` (``pca_fullfr.py:1205-1236`` -> ``pca_grid`` with
    ``scale_list`` and ``initial_4dshape``, ``utils_pca.py:191-228``): ONE decomposition of the rescaled stack with
    max(pclist) components; for every entry of the list the truncated residuals are descaled, collapsed over the
    channels of each ADI frame WITH THE TEMPORAL ``collapse`` MODE (that is what the reference passes to ``scwave``
    there, not ``collapse_ifs``), derotated and combined.  The reference ignores ``cube_ref`` on this branch
    (``cube_ref=None``, :1212).  Returns (cubeout (len(pclist), H, W) device tensor, pclist)."""
    from .pca_fullfr import _pca_grid_device
    z, n, y_in, x_in = cube.shape
    angle_list, scale_list = _check_single_inputs(cube, angle_list, scale_list)
    dev = require_cuda()
    sp = _SinglePass(z, n, y_in, scale_list, ifs_collapse_range, crop_ifs, dev)
    big = sp.rescaled_stack(to_device_f32(cube, dev))
    if verbose:
        print("{} total frames".format(n * z))
        print("Performing single-pass PCA")
    hook = lambda res: sp.descale_collapse(res, collapse)[1]             # noqa: E731
    return _pca_grid_device(big, None, -angle_list, range_pcs, scaling, mask_center_px, svd_mode, collapse,
                            weights=weights, residual_hook=hook, **rot_options)
