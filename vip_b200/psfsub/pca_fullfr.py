"""Full-frame PCA for ADI / RDI cubes on the B200 (drop-in for ``vip_hci.psfsub.pca``).

Reference: ``src/vip_hci/psfsub/pca_fullfr.py`` -- ``PCA_Params`` :93-134, ``pca`` :137-798,
``_adi_rdi_pca`` :801-1035, ``_project_subtract`` :1552-1737.

The call signature, positional-argument order, ``algo_params`` / ``rot_options`` handling,
exceptions and return layouts follow the reference.  The arithmetic runs on the GPU through
``libvipb200.so`` only; option combinations that are not implemented on the GPU raise
``NotImplementedError`` (there is no CPU fallback).
"""
from dataclasses import dataclass
from enum import Enum
from typing import List, Tuple, Union

import numpy as np
import torch

from .. import kernels
from .._device import require_cuda, to_device_f32, to_host
from ..config.paramenum import ALGO_KEY, Adimsdi, Collapse, Imlib, Interpolation, SvdMode
from ..config.utils_conf import check_array
from ..config.utils_param import separate_kwargs_dict, setup_parameters
from ..preproc.derotation import _check_rot_options, derotate_device
from ..preproc.parangles import check_pa_vector
from ..preproc.subsampling import collapse_device
from ..var.shapes import circle_mask
from .svd import Decomposition, _EXACT_MODES, _RAND_MODES, _mode_name, randomized_pcs


@dataclass
class PCA_Params:
    """Parameters of ``pca`` in the reference's declaration order (``pca_fullfr.py:101-134``)."""

    cube: np.ndarray = None
    angle_list: np.ndarray = None
    cube_ref: np.ndarray = None
    scale_list: np.ndarray = None
    ncomp: Union[Tuple, List, float, int] = 1
    svd_mode: Enum = SvdMode.LAPACK
    scaling: Enum = None
    mask_center_px: int = None
    source_xy: Tuple[int] = None
    delta_rot: int = None
    fwhm: float = 4
    adimsdi: Enum = Adimsdi.SINGLE
    crop_ifs: bool = True
    imlib: Enum = Imlib.VIPFFT
    imlib2: Enum = Imlib.VIPFFT
    interpolation: Enum = Interpolation.LANCZOS4
    collapse: Enum = Collapse.MEDIAN
    collapse_ifs: Enum = Collapse.MEAN
    ifs_collapse_range: Union[str, Tuple[int]] = "all"
    smooth: float = None
    smooth_first_pass: float = None
    mask_rdi: np.ndarray = None
    ref_strategy: str = "RDI"
    check_memory: bool = True
    batch: Union[int, float] = None
    nproc: int = 1
    full_output: bool = False
    verbose: bool = True
    weights: np.ndarray = None
    left_eigv: bool = False
    min_frames_pca: int = 10
    max_frames_pca: int = None
    cube_sig: np.ndarray = None
    med_of_npcs: bool = False


def _unsupported(what):
    raise NotImplementedError(
        f"vip_b200.pca: {what} is not implemented on the B200 path yet (no CPU fallback)")


# --------------------------------------------------------------------------------------------------
# device building blocks
# --------------------------------------------------------------------------------------------------


def _in_eps(in_dtype):
    """10 * machine epsilon of the dtype the REFERENCE would scale in (the caller's cube dtype; integer cubes are
    promoted to float64 by ``sklearn.preprocessing.scale``)."""
    dt = np.dtype(in_dtype) if in_dtype is not None else np.dtype(np.float32)
    if not np.issubdtype(dt, np.floating):
        dt = np.dtype(np.float64)
    return 10.0 * float(np.finfo(dt).eps)


def _np_dtype_of(a):
    """numpy dtype of a numpy array or torch tensor (the dtype the reference would compute in)."""
    if isinstance(a, torch.Tensor):
        return np.dtype(str(a.dtype).replace("torch.", ""))
    return np.asarray(a).dtype


def scale_matrix_device(M, scaling, in_dtype=np.float32):
    """``matrix_scaling`` (``var/shapes.py:740-781``): sklearn ``scale`` semantics on the device --
    mean removed along time ('temp-*', axis 0) or space ('spat-*', axis 1); '*-standard' also
    divides by the population std; a std below ``10 * finfo(in_dtype).eps`` (ABSOLUTE threshold, sklearn's
    ``_handle_zeros_in_scale``: about 1.2e-6 for float32 cubes) counts as zero variance and is replaced by 1."""
    if scaling is None:
        return M
    scaling = _mode_name(scaling)
    if scaling not in ("temp-mean", "spat-mean", "temp-standard", "spat-standard"):
        raise ValueError("Scaling mode not recognized")
    axis = 0 if scaling.startswith("temp") else 1
    M64 = M.double()
    mean = M64.mean(dim=axis, keepdim=True)
    out = M64 - mean
    if scaling.endswith("standard"):
        std = torch.sqrt((out * out).mean(dim=axis, keepdim=True))
        std = torch.where(std < _in_eps(in_dtype), torch.ones_like(std), std)
        out = out / std
    return out.to(torch.float32).contiguous()


def prepare_matrix_device(cube_dev, scaling=None, mask_center_px=None, in_dtype=np.float32):
    """``prepare_matrix(mode='fullfr')`` (``var/shapes.py:857-873``): optional central mask, flatten to
    (n, H*W), optional scaling.  Returns a new tensor when anything is modified."""
    n, H, W = cube_dev.shape
    M = cube_dev.reshape(n, H * W)
    if mask_center_px:
        mask = torch.as_tensor(circle_mask((H, W), mask_center_px).reshape(-1)).to(M.device)
        M = M.masked_fill(mask[None, :], 0.0)
    return scale_matrix_device(M, scaling, in_dtype)


def project_subtract_device(cube_dev, ncomp, scaling=None, mask_center_px=None, svd_mode="lapack",
                            cube_ref_dev=None, cube_sig_dev=None, full_output=False, verbose=False,
                            random_state=None, gram=None, pending=None, in_dtype=np.float32):
    """Whole-matrix branch of ``_project_subtract`` (``pca_fullfr.py:1552-1737``) on the device.

    Returns residuals (n,H,W) or (residuals, reconstructed (n,p), V (k,p))."""
    n, H, W = cube_dev.shape
    svd_mode = _mode_name(svd_mode)
    if not isinstance(ncomp, (int, np.integer, float, np.floating)):
        raise TypeError("Type not recognized for ncomp, should be int or float")

    matrix = prepare_matrix_device(cube_dev, scaling, mask_center_px, in_dtype)
    if cube_sig_dev is None:
        matrix_emp = matrix
    else:
        matrix_emp = matrix - cube_sig_dev.reshape(cube_sig_dev.shape[0], -1)
    ref_lib = (prepare_matrix_device(cube_ref_dev, scaling, mask_center_px, in_dtype)
               if cube_ref_dev is not None else matrix_emp)

    dec = None
    if isinstance(ncomp, (float, np.floating)):
        # ncomp given as a cumulative explained variance ratio (pca_fullfr.py:1624-1637); the
        # reference decomposes the *cube* itself for this (SVDecomposer(cube, ...))
        if not 1 > ncomp > 0:
            raise ValueError("if `ncomp` is float, it must lie in the interval (0,1]")
        dec_cube = Decomposition(prepare_matrix_device(cube_dev, scaling, None, in_dtype))
        ncomp = int(np.searchsorted(dec_cube.cevr(), ncomp) + 1)
        if verbose:
            print("Components used : {}".format(ncomp))
        if ref_lib is matrix and not mask_center_px:
            dec = dec_cube
    ncomp = int(ncomp)

    nr, p = ref_lib.shape
    if ncomp > min(nr, p):
        msg = "{} PCs cannot be obtained from a matrix with size [{},{}]."
        msg += " Increase the size of the patches or request less PCs"
        raise RuntimeError(msg.format(ncomp, nr, p))

    # The projection runs in high precision (kernels.project_subtract_hp): fp64 coefficients, the components as an
    # error-free fp32 pair, fp64 accumulation -- fp32 copies of C and V leave a frame-independent error in every
    # residual pixel that survives the temporal median (csrc/proj.cu).  V (returned with full_output) is the hi part.
    if svd_mode in _EXACT_MODES:
        if dec is None:
            dec = Decomposition(ref_lib, ncomp, G=gram if ref_lib is matrix else None, pending=pending)
        V, Vlo = dec.pcs_hilo(ncomp)
        if ref_lib is matrix_emp:
            Cm = dec.coeffs64(ncomp)              # = matrix_emp . V^T from the eigenpairs
        else:
            Cm = kernels.cross_gram(matrix_emp, V) + kernels.cross_gram(matrix_emp, Vlo)
    elif svd_mode in _RAND_MODES:
        rs = random_state
        if rs is None:
            rs = np.random.mtrand._rand
        elif not isinstance(rs, np.random.RandomState):
            rs = np.random.RandomState(rs)
        omega = rs.normal(size=(nr, ncomp + 10))
        (V, Vlo), Cm = randomized_pcs(ref_lib, ncomp, omega, hilo=True, coeffs=True)
        if ref_lib is not matrix_emp:
            Cm = kernels.cross_gram(matrix_emp, V) + kernels.cross_gram(matrix_emp, Vlo)
    else:
        raise ValueError("The SVD `mode` is not recognized")
    if verbose:
        print("Done SVD/PCA on the GPU (vip_b200, svd_mode={})".format(svd_mode))

    residuals = kernels.project_subtract_hp(matrix, Cm, V, Vlo)
    residuals_cube = residuals.reshape(n, H, W)
    if full_output:
        return residuals_cube, kernels.sub(matrix, residuals), V
    return residuals_cube


def sky_pca_residuals_device(cube_dev, ref_dev, masks, ncomp, full_output=False):
    """PCA with data imputation: ``cube_subtract_sky_pca(cube, cube_ref, mask_rdi, ncomp, full_output=True)``
    (``preproc/skysubtraction.py:36-260``) as reached from ``_adi_rdi_pca`` (``pca_fullfr.py:966-972``).

    The reference builds four masked copies of the cubes, the eigen-images of the reference ("sky") cube on the
    ANCHOR mask (all of them, through an SVD of the nr x nr Gramian), projects the anchor-masked science frames on
    them, inverts the nr x nr Gramian of the eigen-images and rebuilds the first ``ncomp`` terms on the BOAT mask.  In
    exact arithmetic that Gramian is diag(lambda), so with (lambda_j, v_j) the leading eigenpairs of
    G = S_a S_a^T (S_a: anchor-masked reference frames) the model of frame i is
    ``sum_{j<ncomp} (v_j . (S_a C_a[i])) / lambda_j * (v_j^T S_b)``: one Gramian, one top-k eigensolve, one cross
    Gramian ``C_a S_a^T``, the boat components ``v_j^T S_b`` as an error-free fp32 pair and the high-precision
    projection kernel of the PCA path -- the sign of v_j cancels.
    Returns residuals (n,H,W) or (residuals, sky_opt (n,H,W), sky_boat_cube (nr,H,W): ALL nr boat components)."""
    n, y, x = cube_dev.shape
    nr = ref_dev.shape[0]
    if tuple(ref_dev.shape[1:]) != (y, x):
        raise TypeError("Science and Sky frames sizes do not match")
    if type(masks) not in (list, tuple):
        mask_anchor = np.asarray(masks)
        mask_boat = np.ones(mask_anchor.shape)
    elif len(masks) != 2:
        raise TypeError("Science and Reference frames sizes do not match")
    else:
        mask_anchor, mask_boat = np.asarray(masks[0]), np.asarray(masks[1])
    if mask_anchor.shape != (y, x) or mask_boat.shape != (y, x):
        raise IndexError("mask_rdi: masks must have the shape of a frame")
    if not isinstance(ncomp, (int, np.integer)):
        raise TypeError("'float' object cannot be interpreted as an integer")          # range(ncomp) in the reference
    k = int(ncomp)
    if k > nr:
        raise IndexError("index {} is out of bounds for axis 0 with size {}".format(nr, nr))
    dev = cube_dev.device
    p = y * x
    za = torch.as_tensor(mask_anchor == 0).reshape(-1).to(dev)
    zb = torch.as_tensor(mask_boat == 0).reshape(-1).to(dev)
    ref2, sci2 = ref_dev.reshape(nr, p), cube_dev.reshape(n, p)
    Sa, Sb = ref2.masked_fill(za[None, :], 0.0), ref2.masked_fill(zb[None, :], 0.0)
    Ca, Cb = sci2.masked_fill(za[None, :], 0.0), sci2.masked_fill(zb[None, :], 0.0)
    dec = Decomposition(Sa, None if full_output else k)          # eigenpairs of S_a S_a^T (all of them for the pcs)
    lam, Vk = dec.evals[:k], dec.U[:k].contiguous()
    Bhi, Blo = kernels.pcs_hilo(Vk, Sb)                          # boat components v_j^T S_b, (k, p)
    T = kernels.cross_gram(Ca, Sa)                               # (n, nr) fp64 = C_a S_a^T
    Cm = ((T @ Vk.t()) / lam[None, :]).contiguous()              # (n, k) coefficients on the anchor
    res = kernels.project_subtract_hp(Cb, Cm, Bhi, Blo)
    if not full_output:
        return res.reshape(n, y, x)
    sky_opt = kernels.sub(Cb, res)
    boat_all = kernels.pcs(dec.U.contiguous(), Sb)               # (nr, p)
    return res.reshape(n, y, x), sky_opt.reshape(n, y, x), boat_all.reshape(nr, y, x)


def _adi_rdi_pca_device(cube, cube_ref, angle_list, ncomp, scaling, mask_center_px, svd_mode, collapse,
                        verbose, full_output, weights=None, cube_sig=None, random_state=None,
                        keep_on_device=False, source_xy=None, delta_rot=None, fwhm=4, min_frames_pca=10,
                        max_frames_pca=None, left_eigv=False, mask_rdi=None, _defer_check=True, **rot_options):
    """``_adi_rdi_pca`` (``pca_fullfr.py:801-1035``) for scalar ``ncomp`` without ``batch``: PCA residuals (whole matrix, or frame by frame with a PA-rejection library when
    ``source_xy`` is given) -> derotation -> collapse."""
    n, y, x = cube.shape
    angle_list = check_pa_vector(np.asarray(angle_list))
    if n != angle_list.shape[0]:
        raise ValueError("`angle_list` vector has wrong length. It must equal the number of frames in the cube")
    if not np.isscalar(ncomp):
        raise TypeError("`ncomp` must be an int, float, tuple or list in the ADI case")
    nref = cube_ref.shape[0] if cube_ref is not None else n
    if isinstance(ncomp, (int, np.integer)) and ncomp > nref:
        ncomp = min(ncomp, nref)
        print("Number of PCs too high (max PCs={}), using {} PCs instead.".format(nref, ncomp))
    elif ncomp <= 0:
        raise ValueError("Number of PCs too low. It should be > 0.")

    mask_val = float(rot_options.get("mask_val", np.nan))
    interp_zeros = bool(rot_options.get("interp_zeros", False))
    _check_rot_options(rot_options.get("imlib", "vip-fft"), rot_options.get("cxy"),
                       rot_options.get("border_mode", "constant"), rot_options.get("edge_blend"), cube.shape)

    dev = require_cuda()
    as_dev = (lambda a: a.to(dev).float()) if isinstance(cube, torch.Tensor) else (lambda a: to_device_f32(a, dev))
    gram = None
    if mask_rdi is not None and cube_ref is None:
        raise TypeError("`mask_rdi` (PCA with data imputation) needs a reference cube (`cube_ref`)")
    plain = (cube_ref is None and cube_sig is None and scaling is None and not mask_center_px
             and source_xy is None
             and _mode_name(svd_mode) in _EXACT_MODES and isinstance(ncomp, (int, np.integer)))
    if (plain and isinstance(cube, np.ndarray) and cube.dtype == np.float32 and cube.flags["C_CONTIGUOUS"]):
        # host cube: upload in pixel slabs and accumulate the Gramian while the next slab is in flight
        M, gram = kernels.upload_and_gram(cube.reshape(n, y * x), dev)
        cube_dev = M.reshape(n, y, x)
    else:
        cube_dev = as_dev(cube)
    ref_dev = as_dev(cube_ref) if cube_ref is not None else None
    sig_dev = as_dev(cube_sig) if cube_sig is not None else None

    pending = None
    if mask_rdi is not None:
        # PCA with data imputation (:966-972): pcs = boat components of the reference cube, recon = the sky model
        out_rdi = sky_pca_residuals_device(cube_dev, ref_dev, mask_rdi, ncomp, full_output)
        residuals_cube, recon, V = out_rdi if full_output else (out_rdi, None, None)
    elif source_xy is not None:
        residuals_cube, recon, nfrslib = pa_rejection_residuals_device(
            cube_dev, ref_dev, sig_dev, angle_list, ncomp, scaling, mask_center_px, svd_mode, source_xy,
            delta_rot, fwhm, min_frames_pca, max_frames_pca, full_output, in_dtype=_np_dtype_of(cube))
        V = None
        if verbose:
            print("Size LIB: min={:.1f} / 10th perc.={:.1f} / median={:.1f} / 90th perc.={:.1f} / max={:.1f}".format(
                np.min(nfrslib), np.percentile(nfrslib, 10), np.median(nfrslib), np.percentile(nfrslib, 90),
                np.max(nfrslib)))
    else:
        # The convergence flag of the subspace eigensolver is checked AFTER the whole pipeline has been
        # enqueued (one synchronisation at the end instead of a pipeline bubble behind the eigensolver: the
        # host would otherwise sit idle for 1.6 ms at config 2 and only then start launching the rest)
        pending = [] if _defer_check else None
        res = project_subtract_device(cube_dev, ncomp, scaling, mask_center_px, svd_mode, ref_dev, sig_dev,
                                      full_output=full_output, verbose=verbose, random_state=random_state,
                                      gram=gram, pending=pending, in_dtype=_np_dtype_of(cube))
        if full_output:
            residuals_cube, recon, V = res
        else:
            residuals_cube = res
    residuals_cube_ = derotate_device(residuals_cube, -angle_list, mask_val=mask_val, interp_zeros=interp_zeros)
    frame = collapse_device(residuals_cube_, mode=collapse, w=weights)
    if mask_center_px:
        mask = torch.as_tensor(circle_mask((y, x), mask_center_px)).to(dev)
        residuals_cube_ = residuals_cube_.masked_fill(mask[None], 0.0)
        frame = frame.masked_fill(mask, 0.0)
    if source_xy is None and mask_rdi is None and pending:
        torch.cuda.current_stream().synchronize()
        if any(int(rec[1]) == 0 for rec in pending):
            # rare: subspace iteration stalled (flat spectrum) -> redo with the synchronous path, which falls
            # back to the full Jacobi solver
            return _adi_rdi_pca_device(cube, cube_ref, angle_list, ncomp, scaling, mask_center_px, svd_mode,
                                       collapse, verbose, full_output, weights=weights, cube_sig=cube_sig,
                                       random_state=random_state, keep_on_device=keep_on_device,
                                       left_eigv=left_eigv, _defer_check=False, **rot_options)
    if verbose:
        print("Done de-rotating and combining")
    if full_output and mask_rdi is not None:
        out = (V, recon, residuals_cube, residuals_cube_, frame)       # pcs = sky_boat_cube, recon = sky_opt (:969-972)
    elif full_output and source_xy is not None:
        out = (recon.reshape(n, y, x), residuals_cube, residuals_cube_, frame)       # pca_fullfr.py:996-1000
    elif full_output and left_eigv:
        # `left_eigv` projects on the temporal (left) singular vectors U_k: U_k U_k^T M = M V_k^T V_k, the same
        # reconstruction as the pixel-space projection (pca_fullfr.py:1697-1700, 1723-1726); only the returned
        # "pcs" differ: V.T of the reference = U_k^T, shape (k, n) (:905).  U_k S_k = M_emp V_k^T.
        US = kernels.cross_gram(V, scale_matrix_device(cube_dev.reshape(n, y * x), scaling,
                                                       _np_dtype_of(cube)))                   # (k, n) = S_k U_k^T
        pcs_left = (US / torch.linalg.vector_norm(US, dim=1, keepdim=True)).to(torch.float32)
        out = (pcs_left, recon.reshape(n, y, x), residuals_cube, residuals_cube_, frame)
    elif full_output:
        out = (V.reshape(V.shape[0], y, x), recon.reshape(n, y, x), residuals_cube, residuals_cube_, frame)
    else:
        out = frame
    return out


def pa_rejection_residuals_device(cube_dev, ref_dev, sig_dev, angle_list, ncomp, scaling, mask_center_px,
                                  svd_mode, source_xy, delta_rot, fwhm, min_frames_pca, max_frames_pca,
                                  full_output=False, in_dtype=np.float32):
    """The ``source_xy`` branch of ``_adi_rdi_pca`` (``pca_fullfr.py:911-965``) with ``_project_subtract``'s
    frame-by-frame mode (:1640-1710): frame f is projected on the PCs of the library of frames whose
    parallactic angle differs from PA_f by more than the threshold set by ``delta_rot`` x ``fwhm`` at the
    separation of ``source_xy`` (+ all reference frames).

    The reference runs n SVDs of (library x pixels) matrices.  Here: ONE Gramian of the stacked matrix,
    then per frame the leading eigenpairs of the library's sub-block:
    ``(lambda, E) = eig_k(G[I,I])``, ``w = E diag(1/lambda) E^T G[I,f]``, residual ``= M_f - w^T M_emp[I]``.
    Libraries of up to 256 frames with ncomp <= 24 go through the batched kernel of the annular path
    (``vb_annular_weights_f64``); larger ones through the full-size eigensolvers, one frame at a time.
    Returns (residuals (n,H,W), reconstruction (n,p) or None, library sizes)."""
    from ..preproc.derotation import _compute_pa_thresh, _find_indices_adi
    from ..var.coords import dist, frame_center
    n, y, x = cube_dev.shape
    dev = cube_dev.device
    svd_mode = _mode_name(svd_mode)
    if svd_mode not in _EXACT_MODES:
        _unsupported(f"svd_mode={svd_mode!r} with `source_xy`")
    if delta_rot is None or fwhm is None:
        raise TypeError("Delta_rot or fwhm parameters missing. Needed forPA-based rejection of frames from the "
                        "library")
    if not isinstance(ncomp, (int, np.integer, float, np.floating)):
        raise TypeError("Type not recognized for ncomp, should be int or float")
    matrix = prepare_matrix_device(cube_dev, scaling, mask_center_px, in_dtype)
    if isinstance(ncomp, (float, np.floating)):
        # CEVR of the whole cube -> ncomp (pca_fullfr.py:1624-1637): SVDecomposer(cube, scaling) sees the cube
        # WITHOUT the central mask
        if not 1 > ncomp > 0:
            raise ValueError("if `ncomp` is float, it must lie in the interval (0,1]")
        unmasked = matrix if not mask_center_px else prepare_matrix_device(cube_dev, scaling, None, in_dtype)
        ncomp = int(np.searchsorted(Decomposition(unmasked, None).cevr(), ncomp) + 1)
    ncomp = int(ncomp)
    matrix_emp = matrix if sig_dev is None else matrix - sig_dev.reshape(sig_dev.shape[0], -1)
    matrix_ref = (prepare_matrix_device(ref_dev, scaling, mask_center_px, in_dtype)
                  if ref_dev is not None else None)
    nref = 0 if matrix_ref is None else matrix_ref.shape[0]

    yc, xc = frame_center((y, x))
    x1, y1 = source_xy
    pa_thr = _compute_pa_thresh(dist(yc, xc, y1, x1), fwhm, delta_rot)
    truncate = max_frames_pca is not None
    lists = [_find_indices_adi(angle_list, f, pa_thr, truncate=truncate, max_frames=max_frames_pca)
             for f in range(n)]
    msg = "{} frames comply to delta_rot condition < less than "
    msg1 = msg + "min_frames_pca ({}). Try decreasing delta_rot or min_frames_pca"
    msg2 = msg + "ncomp ({}). Try decreasing the parameter delta_rot or ncomp"
    nfrslib = []
    for idx in lists:
        L = len(idx) + nref
        if L < min_frames_pca:
            raise RuntimeError(msg1.format(L, min_frames_pca))
        if L < ncomp:
            raise RuntimeError(msg2.format(L, ncomp))
        nfrslib.append(L)
    p = y * x
    if ncomp > p:
        raise RuntimeError("{} PCs cannot be obtained from a matrix with size [{},{}]. Increase the size of the "
                           "patches or request less PCs".format(ncomp, max(nfrslib), p))

    lib = matrix_emp if matrix_ref is None else torch.cat((matrix_emp, matrix_ref))     # library rows after the cube
    ntot = n + nref
    G = kernels.gram(lib)
    Lmax = max(nfrslib)
    ref_rows = np.arange(n, ntot, dtype=np.int32)
    if Lmax <= 256 and ncomp <= 24:
        idx_host = np.zeros((n, Lmax), dtype=np.int32)
        lens = np.asarray(nfrslib, dtype=np.int32)
        for f, idx in enumerate(lists):
            idx_host[f, :len(idx)] = idx
            idx_host[f, len(idx):lens[f]] = ref_rows
        W, iters = kernels.annular_weights(G, torch.from_numpy(idx_host).to(dev), torch.from_numpy(lens).to(dev),
                                           torch.arange(n, dtype=torch.int32, device=dev), ncomp)
        if bool((iters < 0).any()):
            raise RuntimeError("vip_b200.pca: per-frame eigenproblems did not converge")
    else:
        W = torch.zeros((n, ntot), dtype=torch.float32, device=dev)
        for f, idx in enumerate(lists):
            I = torch.from_numpy(np.concatenate((idx, ref_rows)).astype(np.int64)).to(dev)
            Gs = G.index_select(0, I).index_select(1, I).contiguous()
            L = I.numel()
            if kernels.topk_supported(L, ncomp):
                lam, E, info = kernels.eigh_topk(Gs, ncomp)
                if not info["converged"]:
                    lam, E, _ = kernels.eigh(Gs)
            else:
                lam, E, _ = kernels.eigh(Gs)
            lam, E = lam[:ncomp], E[:ncomp]                        # E rows = eigenvectors of G[I,I]
            g = G[f].index_select(0, I)                             # M_emp[f] . M_emp[I]^T
            w = E.t() @ ((E @ g) / lam)
            W[f, I] = w.to(torch.float32)
    recon = None
    if full_output:
        recon = torch.zeros_like(matrix)
        kernels.gemm(W.unsqueeze(0), lib.unsqueeze(0), recon.unsqueeze(0))
    R = matrix.clone()                                   # same bits with and without full_output
    kernels.gemm(W.unsqueeze(0), lib.unsqueeze(0), R.unsqueeze(0), alpha=-1.0, beta=1.0)
    return R.reshape(n, y, x), recon, nfrslib


def _grid_pclist(range_pcs, n):
    """``pca_grid`` PC list (``utils_pca.py:284-305``)."""
    if isinstance(range_pcs, list):
        return list(range_pcs)
    if range_pcs is None:
        pcmin, pcmax, step = 1, n - 1, 1
    elif len(range_pcs) == 2:
        pcmin, pcmax = range_pcs
        pcmax = min(pcmax, n)
        step = 1
    elif len(range_pcs) == 3:
        pcmin, pcmax, step = range_pcs
        pcmax = min(pcmax, n)
    else:
        raise TypeError("`range_pcs` must be None or a tuple, corresponding"
                        "to (PC_INI, PC_MAX) or (PC_INI, PC_MAX, STEP)")
    return list(range(pcmin, pcmax + 1, step))


def _pca_grid_device(cube, cube_ref, rot_angles, range_pcs, scaling, mask_center_px, svd_mode, collapse,
                     weights=None, residual_hook=None, **rot_options):
    """``pca_grid(mode='fullfr', source_xy=None)`` (``utils_pca.py:25-428``) as reached from
    ``_adi_rdi_pca`` (``pca_fullfr.py:1010-1035``): ONE decomposition with max(pclist) components, then
    for every entry of the list the truncated projection/subtraction, derotation and collapse.
    ``residual_hook`` (ADI+mSDI single pass, ``utils_pca.py:201-223``): maps the residual cube of the z*n rescaled
    frames to the (n, H, W) cube that is derotated (descaling + collapse over the channels of each ADI frame).
    Returns (cubeout (len(pclist),H,W) device tensor, pclist)."""
    n, y, x = cube.shape
    rot_angles = np.asarray(rot_angles)          # = -angle_list (cube_derotate convention)
    if _mode_name(svd_mode) not in _EXACT_MODES:
        _unsupported(f"svd_mode={_mode_name(svd_mode)!r} with a tuple/list `ncomp`")
    pclist = _grid_pclist(range_pcs, n)
    pcmax = max(pclist)
    mask_val = float(rot_options.get("mask_val", np.nan))
    interp_zeros = bool(rot_options.get("interp_zeros", False))
    _check_rot_options(rot_options.get("imlib", "vip-fft"), rot_options.get("cxy"),
                       rot_options.get("border_mode", "constant"), rot_options.get("edge_blend"), cube.shape)
    dev = require_cuda()
    in_dtype = _np_dtype_of(cube)
    matrix = prepare_matrix_device(to_device_f32(cube, dev), scaling, mask_center_px, in_dtype)
    ref_lib = (prepare_matrix_device(to_device_f32(cube_ref, dev), scaling, mask_center_px, in_dtype)
               if cube_ref is not None else matrix)
    nr, p = ref_lib.shape
    if pcmax > min(nr, p):
        msg = "{} PCs cannot be obtained from a matrix with size [{},{}]."
        msg += " Increase the size of the patches or request less PCs"
        raise RuntimeError(msg.format(pcmax, nr, p))
    dec = Decomposition(ref_lib, pcmax)
    V, Vlo = dec.pcs_hilo(pcmax)
    if ref_lib is matrix:
        Cm = dec.coeffs64(pcmax)
    else:
        Cm = kernels.cross_gram(matrix, V) + kernels.cross_gram(matrix, Vlo)
    frames = []
    for pc in pclist:
        residuals = kernels.project_subtract_hp(matrix, Cm[:, :pc].contiguous(), V[:pc], Vlo[:pc]).reshape(n, y, x)
        if residual_hook is not None:
            residuals = residual_hook(residuals)
        der = derotate_device(residuals, rot_angles, mask_val=mask_val, interp_zeros=interp_zeros)
        frames.append(collapse_device(der, mode=collapse, w=weights))
    return torch.stack(frames), pclist


def _grid_snr_table(cubeout, pclist, source_xy, fwhm, verbose=False):
    """S/N-optimised number of components (``pca_grid`` with ``source_xy``, ``utils_pca.py:242-275, 352-418``): for
    every frame of the grid the mean S/N and mean aperture flux over the pixels of the FWHM disc around the source
    (fmerit='mean'), all S/N evaluations on the GPU.  Returns (index of the best frame, pandas table)."""
    if fwhm is None:
        raise ValueError("if source_xy is provided, so should fwhm")
    from ..metrics.snr_source import snr_points_device
    from ..var.shapes import disk_indices
    import pandas as pd
    x, y = source_xy
    yy, xx = disk_indices(y, x, fwhm / 2.0, cubeout.shape[1:])
    snrlist, fluxlist = [], []
    for i in range(cubeout.shape[0]):
        sv, fl = snr_points_device(cubeout[i].contiguous(), xx, yy, fwhm)
        snr_value = float(sv.mean().item())
        snrlist.append(0 if np.isnan(snr_value) else snr_value)
        fluxlist.append(float(fl.mean().item()))
    argmax = int(np.argmax(snrlist))
    table = pd.DataFrame({"PCs": pclist, "S/Ns": snrlist, "fluxes": fluxlist})
    if verbose:
        print("Number of steps", len(pclist))
        print("Optimal number of PCs = {}, for S/N={:.3f}".format(pclist[argmax], snrlist[argmax]))
    return argmax, table


def _pca_4d_channels(p, rot_options):
    """4-d input without ``scale_list``: independent ADI/RDI PCA per spectral channel, then
    ``cube_collapse(ifs_adi_frames, collapse_ifs)`` (``pca_fullfr.py:543-658``, returns :762-783)."""
    nch, nz, ny, nx = p.cube.shape
    for name in ("mask_rdi", "source_xy", "batch", "smooth", "cube_sig"):
        if getattr(p, name) is not None:
            _unsupported(f"`{name}` with a 4-d cube")
    if p.left_eigv:
        _unsupported("`left_eigv`")
    if _mode_name(p.imlib) != "vip-fft":
        _unsupported(f"imlib={_mode_name(p.imlib)!r}")
    grid_len = None
    if not isinstance(p.ncomp, list):
        ncomp = [p.ncomp] * nch
    elif len(p.ncomp) != nch:
        grid_len = len(p.ncomp)
        ncomp = [p.ncomp] * nch
    else:
        ncomp = p.ncomp
    if np.isscalar(p.fwhm):
        p.fwhm = [p.fwhm] * nch                 # the reference rewrites the attribute too (:555-556)
    dev = require_cuda()
    pcs, recon, res, res_, frames, pclist = [], [], [], [], [], []
    grid_case = False
    for ch in range(nch):
        cube_ref = None
        if p.cube_ref is not None:
            if p.cube_ref[ch].ndim != 3:
                raise TypeError("Ref cube has wrong format for 4d input cube")
            if p.ref_strategy == "RDI":
                cube_ref = p.cube_ref[ch]
            elif p.ref_strategy == "ARDI":
                cube_ref = np.concatenate((p.cube[ch], p.cube_ref[ch]))
            else:
                raise TypeError("ref_strategy argument not recognized.Should be 'RDI' or 'ARDI'")
        if isinstance(ncomp[ch], (tuple, list)):
            grid_case = True
            cubeout, pcl = _pca_grid_device(p.cube[ch], cube_ref, -check_pa_vector(np.asarray(p.angle_list)),
                                            ncomp[ch], p.scaling, p.mask_center_px, p.svd_mode, p.collapse,
                                            weights=p.weights, **rot_options)
            frames.append(cubeout)
            pclist.append(pcl)
        else:
            out = _adi_rdi_pca_device(p.cube[ch], cube_ref, p.angle_list, ncomp[ch], p.scaling, p.mask_center_px,
                                      p.svd_mode, p.collapse, p.verbose, True, weights=p.weights,
                                      **rot_options)
            pcs.append(out[0]); recon.append(out[1]); res.append(out[2]); res_.append(out[3])
            frames.append(out[4])
    ifs_adi_frames = torch.stack(frames)                     # (nch, H, W) or (nch, npc, H, W)
    if grid_case:
        final = torch.stack([collapse_device(ifs_adi_frames[:, i].contiguous(), mode=p.collapse_ifs)
                             for i in range(ifs_adi_frames.shape[1])])
        return dict(grid=True, final=final, pclist=pclist, ifs=ifs_adi_frames)
    frame = collapse_device(ifs_adi_frames, mode=p.collapse_ifs)
    if len({tuple(t.shape) for t in pcs}) > 1:
        # np.array(pcs) in the reference (:646) for channels with different numbers of components
        raise ValueError("setting an array element with a sequence. The requested array has an inhomogeneous "
                         "shape after 1 dimensions.")
    return dict(grid=False, frame=frame, pcs=torch.stack(pcs), recon=torch.stack(recon), res=torch.stack(res),
                res_=torch.stack(res_), ifs=ifs_adi_frames)


def _to_numpy_like(t, ref_dtype):
    """Device result -> numpy with the dtype the reference would return for a cube of ``ref_dtype``."""
    a = to_host(t)
    if a.dtype == np.float32 and ref_dtype != np.float32 and np.issubdtype(ref_dtype, np.floating):
        a = a.astype(ref_dtype)
    return a


def _pca_adimsdi(p, rot_options):
    """ADI+mSDI branch of ``pca`` (``pca_fullfr.py:497-552``, returns :719-760)."""
    from .sdi import adimsdi_doublepca_device, adimsdi_singlepca_device, adimsdi_singlepca_grid_device
    if p.cube.ndim != 4:
        raise TypeError("Input cube is not a 4d array (required with `scale_list`)")
    for name in ("mask_rdi", "smooth_first_pass", "smooth", "batch"):
        if getattr(p, name) is not None:
            _unsupported(f"`{name}` with ADI+mSDI")
    if p.left_eigv:
        _unsupported("`left_eigv`")
    for lib in (p.imlib, p.imlib2):
        if _mode_name(lib) != "vip-fft":
            _unsupported(f"imlib={_mode_name(lib)!r}")
    adimsdi = _mode_name(p.adimsdi)
    # reference library (:499-510): 'ARSDI' -> the science cube joins the library (single pass: concatenated here)
    cube_ref, ref_strategy = p.cube_ref, "RSDI"
    if cube_ref is not None:
        if np.ndim(cube_ref) != 4:
            raise TypeError("Ref cube has wrong format for 4d input cube")
        if "A" in str(p.ref_strategy):
            ref_strategy = "ARSDI"
            if adimsdi == "single":
                cube_ref = np.concatenate((p.cube, cube_ref), axis=1)
    if adimsdi == "double":
        res_ch, res_der, frame = adimsdi_doublepca_device(
            p.cube, p.angle_list, p.scale_list, p.ncomp, scaling=p.scaling, mask_center_px=p.mask_center_px,
            svd_mode=p.svd_mode, collapse=p.collapse, collapse_ifs=p.collapse_ifs,
            ifs_collapse_range=p.ifs_collapse_range, weights=p.weights, verbose=p.verbose, cube_ref=cube_ref,
            ref_strategy=ref_strategy, source_xy=p.source_xy, delta_rot=p.delta_rot, fwhm=p.fwhm,
            min_frames_pca=p.min_frames_pca, max_frames_pca=p.max_frames_pca, cube_sig=p.cube_sig, **rot_options)
        # the reference's mSDI outputs are float64 (rescaling runs in fp64 there)
        if p.full_output:
            return (to_host(frame).astype(np.float64), to_host(res_ch).astype(np.float64),
                    to_host(res_der).astype(np.float64))
        return to_host(frame).astype(np.float64)
    if adimsdi == "single":
        # `cube_sig` and -- for a scalar ncomp -- `source_xy` are not forwarded to / used by _adimsdi_singlepca in the
        # reference (setup_parameters drops what the function does not take): ignored here as well
        if isinstance(p.ncomp, (tuple, list)):
            # PCA grid on the rescaled stack (:1205-1236, returns :740-751); `cube_ref` is ignored there upstream
            cubeout, pclist = adimsdi_singlepca_grid_device(
                p.cube, p.angle_list, p.scale_list, p.ncomp, scaling=p.scaling, mask_center_px=p.mask_center_px,
                svd_mode=p.svd_mode, collapse=p.collapse, ifs_collapse_range=p.ifs_collapse_range,
                crop_ifs=p.crop_ifs, weights=p.weights, verbose=p.verbose, **rot_options)
            final = to_host(cubeout).astype(np.float64)
            if p.source_xy is not None:
                argmax, table = _grid_snr_table(cubeout, pclist, p.source_xy, p.fwhm, p.verbose)
                frame = final[argmax]
                if p.med_of_npcs:
                    final = np.median(final, axis=0)
                return (final, frame, table) if p.full_output else frame
            if p.med_of_npcs:
                final = np.median(final, axis=0)
            return (final, pclist) if p.full_output else final
        allfr, desc, resadi, frame = adimsdi_singlepca_device(
            p.cube, p.angle_list, p.scale_list, p.ncomp, scaling=p.scaling, mask_center_px=p.mask_center_px,
            svd_mode=p.svd_mode, collapse=p.collapse, collapse_ifs=p.collapse_ifs,
            ifs_collapse_range=p.ifs_collapse_range, crop_ifs=p.crop_ifs, weights=p.weights, verbose=p.verbose,
            cube_ref=cube_ref, **rot_options)
        # reference dtypes: the rescaled cube and every np.zeros buffer are float64; cube_desc_residuals
        # is np.zeros_like(cube) (pca_fullfr.py:1159-1170)
        f64 = lambda t: to_host(t).astype(np.float64)
        if p.full_output:
            return f64(frame), f64(allfr), to_host(desc).astype(p.cube.dtype), f64(resadi)
        return f64(frame)
    raise ValueError(f"ADIMSDI value should only be {Adimsdi.SINGLE} or {Adimsdi.DOUBLE}.")


_MEMCHECK_MIN_BYTES = 8 << 30


def _check_device_memory(p):
    """``check_memory`` (``pca_fullfr.py:438-455``, ``config/mem.py:34-65``) against DEVICE memory: the cube (or the
    reference cube) has to fit in HBM; larger inputs are pointed at ``batch`` (incremental PCA)."""
    if not (p.check_memory and p.batch is None and isinstance(p.cube, np.ndarray)):
        return
    from .. import _device
    input_bytes = p.cube_ref.nbytes if p.cube_ref is not None else p.cube.nbytes
    if input_bytes < _MEMCHECK_MIN_BYTES:        # small inputs always fit: no driver query on the common path
        return
    if input_bytes > _device.free_memory_bytes():
        raise RuntimeError("Input is larger than available device memory. Set check_memory=False to override "
                           "this memory check or set `batch` to run incremental PCA (valid for ADI)")


def pca(*all_args: List, **all_kwargs: dict):
    """Full-frame PCA speckle subtraction: drop-in for ``vip_hci.psfsub.pca``.

    Positional arguments map to :class:`PCA_Params` fields in declaration order; keyword arguments
    that are not fields are the ``rot_options`` forwarded to the derotation; ``algo_params=<obj>``
    (any object carrying the fields) bypasses the parsing (``pca_fullfr.py:398-409``).

    Implemented on the GPU: 3-d ADI / ADI+RDI (``ref_strategy`` 'RDI' or 'ARDI') cubes with scalar
    ``ncomp`` (int, or float = CEVR), every ``svd_mode`` of the reference (the deterministic ones share
    one exact path), ``scaling``, ``mask_center_px``, ``cube_sig``, ``collapse`` modes, ``weights``,
    ``full_output``.  Returns exactly what the reference returns for these cases
    (``pca_fullfr.py:717-798``): ``frame`` or ``(frame, pcs, recon, residuals_cube, residuals_cube_)``.
    """
    class_params, rot_options = separate_kwargs_dict(initial_kwargs=all_kwargs, parent_class=PCA_Params)
    algo_params = rot_options.pop(ALGO_KEY, None)
    if algo_params is None:
        algo_params = PCA_Params(*all_args, **class_params)
    p = algo_params

    if p.mask_center_px and len(rot_options) == 0:
        rot_options["mask_val"] = 0
        rot_options["ker"] = 1
        rot_options["interp_zeros"] = True

    if p.batch is None:
        check_array(p.cube, (3, 4), msg="cube")
    elif not isinstance(p.cube, (str, np.ndarray)):
        raise TypeError("`cube` must be a numpy (3d or 4d) array or a str with the full path on disk")

    if p.left_eigv and (p.batch is not None or p.mask_rdi is not None or p.cube_ref is not None):
        raise NotImplementedError("left_eigv is not compatible with 'mask_rdi' nor 'batch'")

    if p.mask_rdi is not None and p.ref_strategy in ("ARDI", "ARSDI"):
        msg = "mask for data imputation detected. This mode can only run with "
        msg += "a pure RDI strategy, while ref_strategy was set to {}"
        raise TypeError(msg.format(p.ref_strategy))

    if p.batch is not None:
        # incremental PCA in mini-batches (pca_fullfr.py:838-856 -> utils_pca.py:431-614)
        if p.scale_list is not None or (not isinstance(p.cube, str) and p.cube.ndim == 4):
            _unsupported("`batch` with 4-d / ADI+mSDI cubes")
        if p.cube_ref is not None:
            raise ValueError("RDI not compatible with batch mode")
        from .incremental import pca_incremental
        res = pca_incremental(p.cube, p.angle_list, batch=p.batch, ncomp=p.ncomp, collapse=p.collapse,
                              verbose=p.verbose, full_output=p.full_output, weights=p.weights, nproc=p.nproc,
                              imlib=p.imlib, interpolation=p.interpolation, **rot_options)
        if p.full_output:
            frame, _, pcs, medians = res
            return frame, pcs, medians                                                  # pca_fullfr.py:762-763
        return res
    if p.scale_list is not None:
        return _pca_adimsdi(p, rot_options)
    if p.cube.ndim == 4:
        r = _pca_4d_channels(p, rot_options)
        dt = p.cube.dtype
        # the reference gathers the channel frames in a float64 np.zeros buffer (:545): float64 outputs
        ifs = to_host(r["ifs"]).astype(np.float64)
        if r["grid"]:
            final = to_host(r["final"]).astype(np.float64)
            if p.med_of_npcs:
                final = np.median(final, axis=0)
            return (final, r["pclist"], ifs) if p.full_output else final
        if isinstance(p.ncomp, (tuple, list)):
            # reference quirk (:766-790): a per-channel list of scalar ncomp is treated as a "grid" by the
            # return logic although no grid was computed -> the (empty) final_residuals_cube list comes back
            final = np.median([], axis=0) if p.med_of_npcs else []
            return (final, [], ifs) if p.full_output else final
        frame = to_host(r["frame"]).astype(np.float64)
        if p.full_output:
            return (frame, _to_numpy_like(r["pcs"], dt), _to_numpy_like(r["recon"], dt),
                    _to_numpy_like(r["res"], dt), _to_numpy_like(r["res_"], dt), ifs)
        return frame
    if p.left_eigv and (p.mask_center_px or p.source_xy is not None or isinstance(p.ncomp, (tuple, list))
                        or p.cube_sig is not None or _mode_name(p.svd_mode) not in _EXACT_MODES):
        _unsupported("`left_eigv` together with mask_center_px / source_xy / cube_sig / tuple ncomp / "
                     "randomized SVD")
    if p.smooth is not None:
        _unsupported("`smooth`")
    imlib = _mode_name(p.imlib)
    if imlib != "vip-fft":
        _unsupported(f"imlib={imlib!r}")

    # 3-D ADI or ADI+RDI (pca_fullfr.py:661-681)
    cube_ref = p.cube_ref
    if cube_ref is not None:
        if p.ref_strategy == "ARDI":
            cube_ref = np.concatenate((p.cube, cube_ref))
            algo_params.cube_ref = cube_ref        # the reference overwrites the attribute too (:670-672)
        elif p.ref_strategy != "RDI":
            raise TypeError("ref_strategy argument not recognized.Should be 'RDI' or 'ARDI'")
    _check_device_memory(p)

    if isinstance(p.ncomp, (tuple, list)):
        # PCA grid: one residual frame per number of components (pca_fullfr.py:1010-1035, returns :766-790)
        if p.cube_sig is not None:
            _unsupported("`cube_sig` with a tuple/list `ncomp`")
        angs = check_pa_vector(np.asarray(p.angle_list))
        if p.cube.shape[0] != angs.shape[0]:
            raise ValueError("`angle_list` vector has wrong length. It must equal the number of frames in the cube")
        cubeout, pclist = _pca_grid_device(p.cube, cube_ref, -angs, p.ncomp, p.scaling, p.mask_center_px,
                                           p.svd_mode, p.collapse, weights=p.weights, **rot_options)
        if p.source_xy is not None:
            # S/N-optimised number of components; returns (:778) (cube, optimal frame, table) with full_output, the
            # optimal frame otherwise
            argmax, table = _grid_snr_table(cubeout, pclist, p.source_xy, p.fwhm, p.verbose)
            final = _to_numpy_like(cubeout, p.cube.dtype)
            frame = final[argmax]
            if p.med_of_npcs:
                final = np.median(final, axis=0)
            return (final, frame, table) if p.full_output else frame
        final = _to_numpy_like(cubeout, p.cube.dtype)
        if p.med_of_npcs:
            final = np.median(final, axis=0)
        return (final, pclist) if p.full_output else final

    func_params = setup_parameters(params_obj=algo_params, fkt=_adi_rdi_pca_device)
    func_params["cube_ref"] = cube_ref
    res = _adi_rdi_pca_device(**func_params, **rot_options)

    dt = p.cube.dtype
    if p.full_output and p.source_xy is not None:
        recon_cube, residuals_cube, residuals_cube_, frame = res                      # pca_fullfr.py:781
        return (_to_numpy_like(frame, dt), _to_numpy_like(recon_cube, dt), _to_numpy_like(residuals_cube, dt),
                _to_numpy_like(residuals_cube_, dt))
    if p.full_output:
        pcs, recon, residuals_cube, residuals_cube_, frame = res
        return (_to_numpy_like(frame, dt), _to_numpy_like(pcs, dt), _to_numpy_like(recon, dt),
                _to_numpy_like(residuals_cube, dt), _to_numpy_like(residuals_cube_, dt))
    return _to_numpy_like(res, dt)
