"""Annular PCA on the B200: drop-in for ``vip_hci.psfsub.pca_annular`` (3-d ADI / ADI+RDI cubes).

Reference: ``src/vip_hci/psfsub/pca_local.py`` -- ``PCA_ANNULAR_Params`` :39-70, ``pca_annular``
:73-462, ``_pca_adi_rdi`` :594-827, ``do_pca_patch`` :830-909.

The reference runs, per annular segment and per frame, an SVD of that frame's PA-thresholded
library.  Here each segment costs one Gramian (``vb_gram_f32``), one batched kernel that solves all
per-frame eigenproblems on sub-blocks of that Gramian (``vb_annular_weights_f64``) and one skinny
GEMM that applies the resulting weights (``R = A - W . A_lib``).  Library index lists are integer
results of the reference's own selection rule and are computed on the host in numpy, bit-exactly.
"""
from dataclasses import dataclass
from enum import Enum
from typing import List, Tuple, Union

import os
import time

import numpy as np
import torch

from .. import kernels
from .._device import require_cuda, to_device_f32, to_host
from ..config.paramenum import ALGO_KEY, Collapse, Imlib, Interpolation, SvdMode
from ..config.utils_param import separate_kwargs_dict, setup_parameters
from ..preproc.derotation import _check_rot_options, _define_annuli, derotate_device
from ..preproc.parangles import check_pa_vector
from ..preproc.subsampling import collapse_device
from ..var.shapes import get_annulus_segments
from .pca_fullfr import scale_matrix_device
from .svd import Decomposition, _EXACT_MODES, _mode_name


@dataclass
class PCA_ANNULAR_Params:
    """Parameters of ``pca_annular`` in the reference's declaration order (``pca_local.py:43-70``)."""

    cube: np.ndarray = None
    angle_list: np.ndarray = None
    cube_ref: np.ndarray = None
    scale_list: np.ndarray = None
    radius_int: int = 0
    fwhm: float = 4
    asize: float = 4
    n_segments: Union[int, List[int], str] = 1
    delta_rot: Union[float, Tuple[float], List[float]] = (0.1, 1)
    delta_sep: Union[float, Tuple[float], List[float]] = (0.1, 1)
    ncomp: Union[int, Tuple, np.ndarray, str] = 1
    svd_mode: Enum = SvdMode.LAPACK
    nproc: int = 1
    min_frames_lib: int = 2
    max_frames_lib: int = 200
    tol: float = 1e-1
    scaling: Enum = None
    imlib: Enum = Imlib.VIPFFT
    interpolation: Enum = Interpolation.LANCZOS4
    collapse: Enum = Collapse.MEDIAN
    collapse_ifs: Enum = Collapse.MEAN
    ifs_collapse_range: Union[str, Tuple[int]] = "all"
    theta_init: int = 0
    weights: np.ndarray = None
    cube_sig: np.ndarray = None
    full_output: bool = False
    verbose: bool = True
    left_eigv: bool = False


_TIMING = bool(int(os.environ.get("VIP_B200_TIMING", "0")))
_T = {}


def _tick(name, t0):
    """Development aid (VIP_B200_TIMING=1): accumulate synchronised wall time per section."""
    if _TIMING:
        torch.cuda.synchronize()
        _T[name] = _T.get(name, 0.0) + time.perf_counter() - t0
        return time.perf_counter()
    return t0


def _unsupported(what):
    raise NotImplementedError(
        f"vip_b200.pca_annular: {what} is not implemented on the B200 path yet (no CPU fallback)")


def _library_indices_loop(angle_list, pa_thr, max_frames):
    """Frame-by-frame restatement of ``_find_indices_adi`` (kept as the exact fallback and test reference)."""
    n = angle_list.shape[0]
    limit = min(n - 1, max_frames)
    out = []
    for f in range(n):
        d = np.abs(angle_list - angle_list[f])
        close = d[:f] < pa_thr
        index_prev = int(np.argmax(close)) if close.any() else f
        far = d[f:] > pa_thr
        index_foll = f + int(np.argmax(far)) if far.any() else n
        kept = np.concatenate((np.arange(0, index_prev), np.arange(index_foll, n)))
        if len(kept) > limit:
            d_pa = np.abs(angle_list[kept] - angle_list[f])
            kept = np.sort(kept[np.argsort(d_pa)][:limit])
        out.append(kept.astype(np.int32))
    return out


def library_indices(angle_list, pa_thr, max_frames):
    """``_find_indices_adi(angle_list, f, pa_thr, truncate=True, max_frames=...)`` for every frame f
    (``preproc/derotation.py:410-496``); returns a list of int32 arrays.

    Vectorised over frames.  The reference keeps the ``limit`` library frames closest in |dPA| through
    ``np.argsort`` and re-sorts them by index, so only the SET matters: it is unambiguous unless the
    limit-th and (limit+1)-th smallest |dPA| are equal; those rows (and only those) go through the very
    same ``np.argsort`` call as the reference, which reproduces its tie-breaking bit-exactly."""
    ang = np.asarray(angle_list)
    n = ang.shape[0]
    limit = min(n - 1, max_frames)
    D = np.abs(ang[None, :] - ang[:, None])                     # D[f, i] = |ang[i] - ang[f]|
    col = np.arange(n)[None, :]
    row = np.arange(n)[:, None]
    close = (D < pa_thr) & (col < row)
    index_prev = np.where(close.any(1), close.argmax(1), np.arange(n))
    far = (D > pa_thr) & (col >= row)
    index_foll = np.where(far.any(1), far.argmax(1), n)
    kept = (col < index_prev[:, None]) | (col >= index_foll[:, None])
    nkept = kept.sum(1)
    out = [None] * n
    over = nkept > limit
    if over.any():
        rows = np.nonzero(over)[0]
        Dm = np.where(kept[rows], D[rows], np.inf)
        v = np.partition(Dm, limit - 1, axis=1)[:, limit - 1]
        sel = Dm <= v[:, None]
        exact = sel.sum(1) == limit
        cols = np.nonzero(sel[exact])[1].reshape(-1, limit).astype(np.int32)
        for r, c in zip(rows[exact], cols):
            out[r] = c
        for r in rows[~exact]:                                    # tie at the truncation boundary
            k = np.nonzero(kept[r])[0]
            d_pa = np.abs(ang[k] - ang[r])
            out[r] = np.sort(k[np.argsort(d_pa)][:limit]).astype(np.int32)
    for r in np.nonzero(~over)[0]:
        out[r] = np.nonzero(kept[r])[0].astype(np.int32)
    return out


def _segment_residuals(A, A_lib, angle_list, pa_thr, ncomp, min_frames_lib, max_frames_lib, A_ref=None,
                       lists=None, tol=1e-1):
    """Residuals of every frame of one segment matrix ``A`` (n,npx) on the device.

    ``A_lib`` = matrix the libraries are drawn from (A, or A - A_sig); ``A_ref`` = optional RDI rows
    stacked in front of every library (``pca_local.py:880-885``)."""
    n, npx = A.shape
    dev = A.device
    auto = isinstance(ncomp, str)
    if auto and pa_thr == 0:
        lists = [np.arange(n, dtype=np.int32)] * n          # every frame: the whole segment matrix as library
    if pa_thr == 0 and not auto:
        # every frame uses the whole segment matrix as library: one decomposition (pca_local.py:874-878)
        lib = A_lib if A_ref is None else torch.cat((A_ref, A_lib))
        k = min(ncomp, min(lib.shape))
        dec = Decomposition(lib, k)
        V = dec.pcs(k)
        Cm = kernels.cross_gram(A_lib, V).to(torch.float32).contiguous()
        return kernels.project_subtract(A, Cm, V), k

    t0 = time.perf_counter()
    if lists is None:
        lists = library_indices(angle_list, pa_thr, max_frames_lib)
    nref = 0 if A_ref is None else A_ref.shape[0]
    for f, idx in enumerate(lists):
        if len(idx) < min_frames_lib and A_ref is None:
            msg = "Too few frames left in the PCA library. "
            msg += "Accepted indices length ({:.0f}) less than {:.0f}. "
            msg += "Try decreasing either delta_rot or min_frames_lib."
            raise RuntimeError(msg.format(len(idx), min_frames_lib))
    Lmax = max(len(i) for i in lists) + nref
    if Lmax > 256 and auto:
        _unsupported(f"ncomp='auto' with libraries of more than 256 frames (got {Lmax}: max_frames_lib + reference "
                     "frames)")
    idx_host = np.zeros((n, max(Lmax, 1)), dtype=np.int32)
    lens = np.zeros(n, dtype=np.int32)
    for f, idx in enumerate(lists):
        L = len(idx) + nref
        lens[f] = L
        if nref:
            idx_host[f, :nref] = np.arange(nref)
        idx_host[f, nref:L] = idx + nref
    lib = A_lib if A_ref is None else torch.cat((A_ref, A_lib))
    if auto:
        # ncomp='auto' (get_eigenvectors, svd.py:622-672): per frame, components are added until the pixel noise of
        # the library residuals decays by less than `tol`; evaluated inside the direct eigen-kernel from the
        # eigenpairs of the library Gramian and the row sums of the library matrix
        G = kernels.gram(lib)
        t0 = _tick("gram", t0)
        W, used = kernels.annular_weights_auto(
            G, torch.from_numpy(idx_host).to(dev), torch.from_numpy(lens).to(dev),
            torch.arange(nref, nref + n, dtype=torch.int32, device=dev), lib.double().sum(dim=1), npx, float(tol))
        if bool((used < 0).any()):
            _unsupported("ncomp='auto' choosing more than 24 principal components (decrease `tol`-sensitivity: "
                         "the noise-decay rule did not stop within 24 components)")
        t0 = _tick("weights (batched eigenproblems, auto)", t0)
        R = A.clone()
        kernels.gemm(W.unsqueeze(0), lib.unsqueeze(0), R.unsqueeze(0), alpha=-1.0, beta=1.0)
        return R, used
    k = min(ncomp, npx)                      # get_eigenvectors clamps to min(shape) (svd.py:694)
    G = kernels.gram(lib)
    t0 = _tick("gram", t0)
    if k > 24 or Lmax > 256:
        # outside the batched kernel's limits (24 components, 256-frame libraries): the same weights
        # w = E diag(1/lambda) E^T G[I, f] frame by frame through the full-size eigensolvers (functional, not fast:
        # ~1 ms per frame and segment); get_eigenvectors clamps ncomp to the library size per frame (svd.py:694)
        W = torch.zeros((n, lib.shape[0]), dtype=torch.float32, device=dev)
        for f in range(n):
            L = int(lens[f])
            I = torch.from_numpy(idx_host[f, :L].astype(np.int64)).to(dev)
            kk = min(k, L)
            Gs = G.index_select(0, I).index_select(1, I).contiguous()
            if kernels.topk_supported(L, kk):
                lam, E, info = kernels.eigh_topk(Gs, kk)
                if not info["converged"]:
                    lam, E, _ = kernels.eigh(Gs)
            else:
                lam, E, _ = kernels.eigh(Gs)
            lam, E = lam[:kk], E[:kk]                             # E rows = eigenvectors of G[I, I]
            g = G[nref + f].index_select(0, I)                     # target (emptied) frame . library rows
            W[f, I] = (E.t() @ ((E @ g) / lam)).to(torch.float32)
        t0 = _tick("weights (frame-by-frame eigenproblems)", t0)
        R = A.clone()
        kernels.gemm(W.unsqueeze(0), lib.unsqueeze(0), R.unsqueeze(0), alpha=-1.0, beta=1.0)
        _tick("apply W.A_lib and subtract", t0)
        return R, k
    W, iters = kernels.annular_weights(
        G, torch.from_numpy(idx_host).to(dev), torch.from_numpy(lens).to(dev),
        torch.arange(nref, nref + n, dtype=torch.int32, device=dev), k)
    if bool((iters < 0).any()):                  # cannot happen: the direct solver always returns
        bad = int((iters < 0).sum())
        raise RuntimeError(f"vip_b200.pca_annular: {bad} per-frame eigenproblems did not converge")
    t0 = _tick("weights (batched eigenproblems)", t0)
    # R = A - W . A_lib  (the per-frame PSF models) as one fp32 GEMM with alpha = -1, beta = 1
    R = A.clone()
    kernels.gemm(W.unsqueeze(0), lib.unsqueeze(0), R.unsqueeze(0), alpha=-1.0, beta=1.0)
    _tick("apply W.A_lib and subtract", t0)
    return R, k


def _pca_adi_rdi_device(cube, angle_list, radius_int=0, fwhm=4, asize=2, n_segments=1, delta_rot=1, ncomp=1,
                        svd_mode="lapack", nproc=None, min_frames_lib=2, max_frames_lib=200, tol=1e-1,
                        scaling=None, imlib="vip-fft", interpolation="lanczos4", collapse="median",
                        full_output=False, verbose=1, cube_ref=None, theta_init=0, weights=None,
                        cube_sig=None, left_eigv=False, **rot_options):
    """``_pca_adi_rdi`` (``pca_local.py:594-827``) with the arithmetic on the GPU."""
    array = cube
    if array.ndim != 3:
        raise TypeError("Input array is not a cube or 3d array")
    if array.shape[0] != angle_list.shape[0]:
        raise TypeError("Input vector or parallactic angles has wrong length")
    n, y, x = array.shape
    angle_list = check_pa_vector(angle_list)
    n_annuli = int((y / 2 - radius_int) / asize)

    if isinstance(delta_rot, tuple):
        delta_rot = np.linspace(delta_rot[0], delta_rot[1], num=n_annuli)
    elif np.isscalar(delta_rot):
        delta_rot = [delta_rot] * n_annuli
    elif len(delta_rot) != n_annuli:
        raise TypeError("If delta_rot is a list it should have n_annuli elements.")

    if isinstance(n_segments, int):
        n_segments = [n_segments for _ in range(n_annuli)]
    elif n_segments == "auto":
        n_segments = [2, 3]
        ld = 2 * np.tan(360 / 4 / 2) * asize
        for i in range(2, n_annuli):
            ang = np.rad2deg(2 * np.arctan(ld / (2 * i * asize)))
            n_segments.append(int(np.ceil(360 / ang)))

    if verbose:
        print("N annuli = {}, FWHM = {:.3f}".format(n_annuli, fwhm))
        print("PCA per annulus (or annular sectors):")

    if _mode_name(svd_mode) not in _EXACT_MODES:
        _unsupported(f"svd_mode={_mode_name(svd_mode)!r}")
    if isinstance(ncomp, list):
        # one residual cube per number of components (pca_local.py:665-668, 799-807).  do_pca_patch
        # truncates the SVD of max(ncomp) components (:893-903), which equals one run per value.
        kw = dict(radius_int=radius_int, fwhm=fwhm, asize=asize, n_segments=n_segments, delta_rot=delta_rot,
                  svd_mode=svd_mode, nproc=nproc, min_frames_lib=min_frames_lib, max_frames_lib=max_frames_lib,
                  tol=tol, scaling=scaling, imlib=imlib, interpolation=interpolation, collapse=collapse,
                  full_output=True, verbose=verbose, cube_ref=cube_ref, theta_init=theta_init, weights=weights,
                  cube_sig=cube_sig, left_eigv=left_eigv)
        runs = [_pca_adi_rdi_device(cube, angle_list, ncomp=int(v), **kw, **rot_options) for v in ncomp]
        cube_out = torch.stack([r[0] for r in runs])
        cube_der = torch.stack([r[1] for r in runs])
        frames = [r[2] for r in runs]
        if full_output:
            return cube_out, cube_der, frames
        return frames
    if isinstance(ncomp, str) and ncomp != "auto":
        raise TypeError("`ncomp` must be an int, a tuple/array of ints, a list or 'auto'")
    _check_rot_options(imlib, rot_options.get("cxy"), rot_options.get("border_mode", "constant"),
                       rot_options.get("edge_blend"), array.shape)

    t0 = time.perf_counter()
    dev = require_cuda()
    if isinstance(array, torch.Tensor):              # stage-1 output of the ADI+mSDI branch: already on the device
        cube_dev = array.to(torch.float32).contiguous().reshape(n, y * x)
    else:
        cube_dev = to_device_f32(array, dev).reshape(n, y * x)
    if isinstance(cube_ref, torch.Tensor):
        ref_dev = cube_ref.to(torch.float32).contiguous().reshape(cube_ref.shape[0], y * x)
    else:
        ref_dev = to_device_f32(cube_ref, dev).reshape(cube_ref.shape[0], y * x) if cube_ref is not None else None
    sig_dev = to_device_f32(cube_sig, dev).reshape(n, y * x) if cube_sig is not None else None
    cube_out = torch.zeros_like(cube_dev)
    t0 = _tick("upload", t0)
    G_full = M_full = None
    if left_eigv:
        # `left_eigv` (pca_local.py:704-707, 755-779): every segment is projected on the leading TEMPORAL singular
        # vectors of the pixels outside it.  Their Gramian is the Gramian of the whole frame minus the Gramian of the
        # segment -- one full-frame product (tensor cores) serves all segments -- whenever the scaling acts per pixel
        # (none, temp-*); the per-frame scalings (spat-*) depend on the pixel set and take the outside matrix itself.
        sc = None if scaling is None else _mode_name(scaling)
        if sc is None or sc.startswith("temp"):
            M_full = scale_matrix_device(cube_dev, scaling)
            G_full = kernels.gram(M_full)

    verbose_ann = (int(verbose) + int(cube_ref is None)) if verbose else verbose
    # library index lists of every annulus (host integer logic, independent of the pixel data): computed
    # up-front on a few host threads so that they do not serialise with the GPU work of each annulus
    thresholds = [_define_annuli(angle_list, ann, n_annuli, fwhm, radius_int, asize, delta_rot[ann],
                                 n_segments[ann], False, True)[0] for ann in range(n_annuli)]
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(8, max(1, n_annuli))) as pool:
        all_lists = list(pool.map(lambda thr: library_indices(angle_list, thr, max_frames_lib) if thr != 0 else None,
                                  thresholds))
    t0 = _tick("index lists (host)", t0)
    for ann in range(n_annuli):
        if isinstance(ncomp, (tuple, np.ndarray)):
            if len(ncomp) != n_annuli:
                raise TypeError("If `ncomp` is a tuple, its length must match the number of annuli")
            ncompann = ncomp[ann] if isinstance(ncomp[ann], str) else int(ncomp[ann])
        else:
            ncompann = ncomp
        pa_thr, inner_radius, _ = _define_annuli(angle_list, ann, n_annuli, fwhm, radius_int, asize,
                                                 delta_rot[ann], n_segments[ann], verbose_ann, True)
        segments = get_annulus_segments((y, x), inner_radius, asize, n_segments[ann], theta_init)
        for yy, xx in segments:
            if yy.size == 0:
                continue
            cols = torch.from_numpy((yy * x + xx).astype(np.int32)).to(dev)
            if left_eigv:
                if G_full is not None:
                    A = kernels.gather_columns(M_full, cols)
                    G_out = G_full - kernels.gram(A)
                    p_out = y * x - cols.numel()
                else:
                    outside = np.ones(y * x, dtype=bool)
                    outside[yy * x + xx] = False
                    cols_out = torch.from_numpy(np.nonzero(outside)[0].astype(np.int32)).to(dev)
                    A = scale_matrix_device(kernels.gather_columns(cube_dev, cols), scaling)
                    G_out = kernels.gram(scale_matrix_device(kernels.gather_columns(cube_dev, cols_out), scaling))
                    p_out = cols_out.numel()
                kk = min(int(ncompann), n, p_out)
                if kernels.topk_supported(n, kk):
                    _, U, info = kernels.eigh_topk(G_out, kk)
                    if not info["converged"]:
                        U = kernels.eigh(G_out)[1][:kk]
                else:
                    U = kernels.eigh(G_out)[1][:kk]
                V = kernels.pcs(U, A)                                            # (kk, npx) = U A
                R = kernels.project_subtract(A, U.t().to(torch.float32).contiguous(), V)
                kernels.scatter_columns(R, cols, cube_out)
                continue
            A = scale_matrix_device(kernels.gather_columns(cube_dev, cols), scaling)
            A_ref = (scale_matrix_device(kernels.gather_columns(ref_dev, cols), scaling)
                     if ref_dev is not None else None)
            A_lib = A - kernels.gather_columns(sig_dev, cols) if sig_dev is not None else A
            R, _ = _segment_residuals(A, A_lib, angle_list, pa_thr, ncompann, min_frames_lib,
                                      max_frames_lib, A_ref, lists=all_lists[ann], tol=tol)
            kernels.scatter_columns(R, cols, cube_out)
        if verbose == 1:
            print("Done PCA with {} for current annulus".format(_mode_name(svd_mode)))

    t0 = time.perf_counter()
    cube_out = cube_out.reshape(n, y, x)
    mask_val = float(rot_options.get("mask_val", np.nan))
    interp_zeros = bool(rot_options.get("interp_zeros", False))
    cube_der = derotate_device(cube_out, -angle_list, mask_val=mask_val, interp_zeros=interp_zeros)
    frame = collapse_device(cube_der, mode=collapse, w=weights)
    _tick("derotate + collapse", t0)
    if _TIMING:
        print("vip_b200.pca_annular timing:", {k_: round(v_, 4) for k_, v_ in _T.items()})
        _T.clear()
    if verbose:
        print("Done derotating and combining.")
    if full_output:
        return cube_out, cube_der, frame
    return frame


def _sdi_library_indices(scal, ann_center, fwhm, delta_sep):
    """``_find_indices_sdi(scal, ann_center, j, fwhm, delta_sep)`` for every channel j (``rescaling.py:916-988``)."""
    from ..preproc.rescaling import _find_indices_sdi
    return [np.asarray(_find_indices_sdi(scal, ann_center, j, fwhm, delta_sep), dtype=np.int32)
            for j in range(len(scal))]


def _pca_sdi_frames_device(cube_dev, scal, radius_int, fwhm, asize, n_segments, delta_sep, ncomp, svd_mode,
                           scaling, collapse_ifs, ifs_collapse_range, theta_init):
    """``_pca_sdi_fr`` (``pca_local.py:470-591``) for ALL ADI frames of a (z, n, H, W) device cube -> (n, H, W).

    Per multi-spectral frame the reference rescales the z channels, then per annular segment and per channel j runs a
    PCA of channel j against the channels that moved radially by at least ``delta_sep`` FWHM
    (``_find_indices_sdi``).  The channel libraries depend on the annulus only, not on the ADI frame: the segment
    matrices of a group of ADI frames are stacked ((F z) x npx), ONE Gramian serves the group (its diagonal z x z
    blocks are the per-frame Gramians), the batched eigen-kernel solves the F z library problems on index lists
    offset into the blocks, and one GEMM applies the weights."""
    from ..preproc.rescaling import check_scal_vector
    from .sdi import RescaleOps, _CHUNK_BYTES
    z, n, H, W = cube_dev.shape
    dev = cube_dev.device
    if H != W:
        raise ValueError("FFT scaling only supports square input arrays")
    scal = np.asarray(scal)
    scale_list = check_scal_vector(scal)
    ops = RescaleOps(scale_list, H, dev)
    S = ops.big
    fwhm = int(np.round(np.mean(fwhm)))
    n_annuli = int((H / 2 - radius_int) / asize)
    if isinstance(n_segments, int):
        n_segments = [n_segments for _ in range(n_annuli)]
    elif isinstance(n_segments, str) and n_segments == "auto":
        n_segments = [2, 3]
        ld = 2 * np.tan(360 / 4 / 2) * asize
        for i in range(2, n_annuli):
            ang = np.rad2deg(2 * np.arctan(ld / (2 * i * asize)))
            n_segments.append(int(np.ceil(360 / ang)))
    if isinstance(delta_sep, (tuple, list)):
        delta_sep_vec = np.linspace(delta_sep[0], delta_sep[1], n_annuli)
    elif np.isscalar(delta_sep):
        delta_sep_vec = [delta_sep] * n_annuli
    else:
        if len(delta_sep) != n_annuli:
            raise TypeError("If delta_sep is a list it should have n_annuli elements.")
        delta_sep_vec = delta_sep
    if _mode_name(svd_mode) not in _EXACT_MODES:
        _unsupported(f"svd_mode={_mode_name(svd_mode)!r}")
    if isinstance(ncomp, str):
        _unsupported("ncomp='auto' in the spectral pass")
    if min(int(ncomp), z) > 24:
        _unsupported("more than 24 principal components per annulus")
    i0, i1 = (0, z) if ifs_collapse_range == "all" else (int(ifs_collapse_range[0]), int(ifs_collapse_range[1]))

    # geometry and channel libraries of every segment (host integer logic, shared by all ADI frames)
    plan = []
    for ann in range(n_annuli):
        inner_radius = radius_int + (ann * asize - 1) if ann == n_annuli - 1 else radius_int + ann * asize
        ann_center = inner_radius + (asize / 2)
        lists = _sdi_library_indices(scal, ann_center, fwhm, delta_sep_vec[ann])
        for yy, xx in get_annulus_segments((S, S), inner_radius, asize, n_segments[ann], theta_init):
            if yy.size:
                plan.append((torch.from_numpy((yy * S + xx).astype(np.int32)).to(dev), lists))

    group = max(1, 256 // z)                                   # ADI frames per Gramian: (group z) <= 256 rows
    per_frame = 4 * z * S * S * 4 * 2
    chunk = max(1, min(n, int(_CHUNK_BYTES // per_frame)))
    out = torch.empty((n, ops.out, ops.out), dtype=torch.float32, device=dev)
    for f0 in range(0, n, chunk):
        f1 = min(n, f0 + chunk)
        F = f1 - f0
        ms = cube_dev[:, f0:f1].permute(1, 0, 2, 3).contiguous()               # (F, z, H, W)
        if ops.pad:
            ms = torch.nn.functional.pad(ms, (ops.pad,) * 4, mode="reflect")
        resc = RescaleOps.apply(ms.reshape(F * z, S, S), ops.Wf, z).reshape(F * z, S * S)
        res = torch.zeros_like(resc)
        for cols, lists in plan:
            npx = cols.numel()
            A = kernels.gather_columns(resc, cols)                              # (F z, npx)
            if scaling is not None:
                A = torch.cat([scale_matrix_device(A[f * z:(f + 1) * z], scaling) for f in range(F)])
            Lmax = max(len(l) for l in lists)
            k = min(int(ncomp), npx)
            for g0 in range(0, F, group):
                g1 = min(F, g0 + group)
                Fg = g1 - g0
                Ag = A[g0 * z:g1 * z]
                idx_host = np.zeros((Fg * z, Lmax), dtype=np.int32)
                lens = np.zeros(Fg * z, dtype=np.int32)
                for f in range(Fg):
                    for j, l in enumerate(lists):
                        idx_host[f * z + j, :len(l)] = l + f * z
                        lens[f * z + j] = len(l)
                G = kernels.gram(Ag)
                Wt, iters = kernels.annular_weights(
                    G, torch.from_numpy(idx_host).to(dev), torch.from_numpy(lens).to(dev),
                    torch.arange(Fg * z, dtype=torch.int32, device=dev), k)
                if bool((iters < 0).any()):
                    raise RuntimeError("vip_b200.pca_annular: spectral eigenproblems did not converge")
                R = Ag.clone()
                # Wt is block diagonal (libraries never leave their frame): apply the Fg diagonal z x z blocks as a batch
                P = Wt.reshape(Fg, z, Fg, z).diagonal(dim1=0, dim2=2).permute(2, 0, 1).contiguous()
                kernels.gemm(P, Ag.reshape(Fg, z, npx), R.reshape(Fg, z, npx), alpha=-1.0, beta=1.0)
                kernels.scatter_columns(R, cols, res[g0 * z:g1 * z])
        desc = RescaleOps.apply(res.reshape(F * z, S, S), ops.Wi, z).reshape(F, z, ops.out, ops.out)
        for f in range(F):
            out[f0 + f] = collapse_device(desc[f, i0:i1], collapse_ifs).float()
    return out


def _pca_annular_adimsdi(p, rot_options):
    """ADI+mSDI branch of ``pca_annular`` (``pca_local.py:332-462``): spectral pass per ADI frame with ``ncomp[0]``,
    then the annular ADI pass with ``ncomp[1]`` (or derotation + collapse only when it is None)."""
    cube = p.cube
    z, n, y_in, x_in = cube.shape
    fwhm = int(np.round(np.mean(p.fwhm)))
    scale_list = np.asarray(p.scale_list)
    if scale_list.ndim > 1:
        raise ValueError("Scaling factors vector is not 1d")
    if not scale_list.shape[0] == z:
        raise ValueError("Scaling factors vector has wrong length")
    if not isinstance(p.ncomp, tuple):
        raise TypeError("`ncomp` must be a tuple of two integers when `cube` is a 4d array")
    ncomp1, ncomp2 = p.ncomp[0], p.ncomp[1]
    if p.cube_ref is not None and (not isinstance(p.cube_ref, np.ndarray) or p.cube_ref.ndim != 4 or
                                   p.cube_ref.shape[0] != z or p.cube_ref.shape[2:] != cube.shape[2:]):
        raise TypeError("Ref cube has wrong format for 4d input cube")
    _check_rot_options(p.imlib, rot_options.get("cxy"), rot_options.get("border_mode", "constant"),
                       rot_options.get("edge_blend"), cube.shape[1:])
    dev = require_cuda()
    cube_dev = to_device_f32(cube, dev)
    if p.verbose:
        print("First PCA subtraction exploiting the spectral variability")
        print("{} spectral channels per IFS frame".format(z))
        print("N annuli = {}, mean FWHM = {:.3f}".format(int((y_in / 2 - p.radius_int) / p.asize), fwhm))
    res_channels = _pca_sdi_frames_device(cube_dev, scale_list, p.radius_int, fwhm, p.asize, p.n_segments,
                                          p.delta_sep, ncomp1, p.svd_mode, p.scaling, p.collapse_ifs,
                                          p.ifs_collapse_range, p.theta_init)
    del cube_dev
    angle_list = np.asarray(p.angle_list)
    if ncomp2 is None:
        if p.verbose:
            print("Skipping the second PCA subtraction")
        mask_val = float(rot_options.get("mask_val", np.nan))
        interp_zeros = bool(rot_options.get("interp_zeros", False))
        cube_out = res_channels
        # the reference passes the angles as they are here (no check_pa_vector; cube_derotate negates)
        cube_der = derotate_device(cube_out, -angle_list, mask_val=mask_val, interp_zeros=interp_zeros)
        frame = collapse_device(cube_der, mode=p.collapse, w=p.weights)
        return cube_out, cube_der, frame
    if p.verbose:
        print("Second PCA subtraction exploiting angular variability")
    res_ref = None
    if p.cube_ref is not None:
        # the same spectral pass on the reference cube (pca_local.py:407-432); its residual frames become the
        # reference library of the ADI pass
        if p.verbose:
            print("First PCA subtraction (spectral) on REF cube")
        res_ref = _pca_sdi_frames_device(to_device_f32(p.cube_ref, dev), scale_list, p.radius_int, fwhm, p.asize,
                                         p.n_segments, p.delta_sep, ncomp1, p.svd_mode, p.scaling, p.collapse_ifs,
                                         p.ifs_collapse_range, p.theta_init)
    func_params = setup_parameters(params_obj=p, fkt=_pca_adi_rdi_device, cube=res_channels, ncomp=ncomp2,
                                   fwhm=fwhm, cube_ref=res_ref, full_output=True)
    return _pca_adi_rdi_device(**func_params, **rot_options)


def pca_annular(*all_args: List, **all_kwargs: dict):
    """Annular (per-frame PA-thresholded library) PCA: drop-in for ``vip_hci.psfsub.pca_annular``.

    Positional arguments map to :class:`PCA_ANNULAR_Params` fields in declaration order; keyword
    arguments that are not fields become ``rot_options``; ``algo_params=<obj>`` bypasses parsing.
    Implemented on the GPU: 3-d ADI and ADI+RDI cubes, int / per-annulus-tuple ``ncomp`` (<= 24),
    ``n_segments`` (int, list or 'auto'), ``delta_rot``, ``scaling``, ``cube_sig``, ``radius_int``,
    every deterministic ``svd_mode``.  Returns ``frame`` or ``(cube_out, cube_der, frame)``.
    """
    class_params, rot_options = separate_kwargs_dict(initial_kwargs=all_kwargs,
                                                     parent_class=PCA_ANNULAR_Params)
    algo_params = rot_options.pop(ALGO_KEY, None)
    if algo_params is None:
        algo_params = PCA_ANNULAR_Params(*all_args, **class_params)
    p = algo_params

    if p.radius_int and len(rot_options) == 0:
        rot_options["mask_val"] = 0
        rot_options["ker"] = 1
        rot_options["interp_zeros"] = True

    if p.left_eigv and (p.cube_ref is not None or p.cube_sig is not None or p.ncomp == "auto"):
        raise NotImplementedError("left_eigv is not compatiblewith 'cube_ref', 'cube_sig', ncomp='auto'")

    if not isinstance(p.cube, np.ndarray):
        raise TypeError("Input array is not a cube or 3d array")
    if p.cube.ndim not in (3, 4):
        raise TypeError("Input array is not a 4d or 3d array")
    dt = p.cube.dtype

    def host(t):
        a = to_host(t)
        if a.dtype == np.float32 and dt != np.float32 and np.issubdtype(dt, np.floating):
            a = a.astype(dt)
        return a

    if p.cube.ndim == 4 and p.scale_list is not None:
        cube_out, cube_der, frame = _pca_annular_adimsdi(p, rot_options)
        if p.full_output:
            return host(cube_out), host(cube_der), host(frame)
        return host(frame)

    if p.cube.ndim == 4:
        # 4-d cube without mSDI: annular ADI/RDI per spectral channel, channel frames combined with
        # `collapse_ifs` (pca_local.py:280-330); the reference rewrites ncomp / fwhm into per-channel lists
        nch = p.cube.shape[0]
        if not isinstance(p.ncomp, list) or len(p.ncomp) != nch:
            p.ncomp = [p.ncomp] * nch
        if np.isscalar(p.fwhm):
            p.fwhm = [p.fwhm] * nch
        outs, ders, frames = [], [], []
        for ch in range(nch):
            cube_ref = None
            if p.cube_ref is not None:
                if p.cube_ref[ch].ndim != 3:
                    raise TypeError("Ref cube has wrong format for 4d input cube")
                cube_ref = p.cube_ref[ch]
            func_params = setup_parameters(params_obj=p, fkt=_pca_adi_rdi_device, cube=p.cube[ch],
                                           fwhm=p.fwhm[ch], ncomp=p.ncomp[ch], full_output=True, cube_ref=cube_ref)
            co, cd, fr = _pca_adi_rdi_device(**func_params, **rot_options)
            outs.append(co); ders.append(cd); frames.append(fr)
        ifs = torch.stack(frames)
        # the channel frames live in a float64 np.zeros buffer in the reference (:282): float64 frame
        if p.collapse_ifs is not None:
            frame = to_host(collapse_device(ifs, mode=p.collapse_ifs)).astype(np.float64)
        else:
            frame = to_host(ifs).astype(np.float64)
        if p.full_output:
            return host(torch.stack(outs)), host(torch.stack(ders)), frame
        return frame

    func_params = setup_parameters(params_obj=p, fkt=_pca_adi_rdi_device, full_output=True)
    cube_out, cube_der, frame = _pca_adi_rdi_device(**func_params, **rot_options)
    if isinstance(frame, list):
        # list `ncomp`: the reference allocates these buffers with np.zeros (float64) and returns a list of frames
        frames = [to_host(f).astype(np.float64) for f in frame]
        if p.full_output:
            return to_host(cube_out).astype(np.float64), to_host(cube_der).astype(np.float64), frames
        return frames
    if p.full_output:
        return host(cube_out), host(cube_der), host(frame)
    return host(frame)
