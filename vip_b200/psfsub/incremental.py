"""Incremental (mini-batch) full-frame ADI PCA: ``pca(..., batch=...)`` -> ``pca_incremental``
(``vip_hci/psfsub/utils_pca.py:431-614``), streamed through the GPU batch by batch.

The reference fits scikit-learn's ``IncrementalPCA`` (``partial_fit`` per batch), then makes a second pass over the
cube: every batch is centred with the model mean, projected on the PCs, subtracted, derotated and collapsed, and the
final frame is the MEDIAN of the per-batch frames.  Here the cube stays on the host (it may be larger than HBM) and
each batch is uploaded twice; per batch the work is the kernels of the in-core path:

* ``partial_fit``: the column sums of the batch (fp64, ``vb_collapse_f32`` weighted sum), the incremental mean
  update, and the SVD of the stacked matrix  [diag(s) V ; X - batch mean ; mean correction]  (k + b + 1 rows) through
  its fp64 Gramian (``vb_gram_f32`` / tcgen05), the fp64 eigensolver and ``vb_pcs_f32`` -- the route of the exact
  in-core PCA.  scikit-learn runs LAPACK on the same stack, in fp32 for the first batch and fp64 afterwards.
* second pass: two rank-1 subtractions with the high and low fp32 parts of the fp64 model mean (an fp64-quality
  centring without an fp64 copy of the batch), ``vb_cross_gram_f32`` -> ``vb_project_subtract_f32`` ->
  ``vb_derotate_f32`` -> ``vb_collapse_f32``.

Outputs are float64 like the reference's (its second pass runs on ``cube - mean_`` with a float64 mean).
"""
import numpy as np
import torch

from .. import kernels
from .. import _device
from ..preproc.derotation import derotate_device, _check_rot_options
from ..preproc.parangles import check_pa_vector
from ..preproc.subsampling import collapse_device
from .svd import Decomposition


def batch_ranges(n_frames, batch, cube_nbytes, available_bytes):
    """Frame ranges of the mini-batches (``utils_pca.py:531-552``): an int is the batch size in frames, a float
    in (0, 1) the fraction of the available memory one batch may take (device memory here: the batches are what
    has to fit in HBM)."""
    if isinstance(batch, (int, np.integer)) and not isinstance(batch, bool):
        size = int(batch)
    elif isinstance(batch, float):
        if not 0 < batch < 1:
            raise ValueError("a float `batch` must lie in (0, 1)")
        size = min(int(n_frames * (batch * available_bytes) / cube_nbytes), n_frames)
    else:
        raise TypeError("`batch` must be an int or float")
    if size < 1:
        raise ValueError("`batch` leaves less than one frame per mini-batch")
    return [(i, min(n_frames, i + size)) for i in range(0, n_frames, size)]


def _column_sums(X):
    """(p,) fp64 sums over the frames of an (b, p) fp32 batch: the fp64-accumulating weighted-sum kernel."""
    return kernels.collapse(X, "wmean", w=np.ones(X.shape[0], dtype=np.float64))


def _subtract_row_f64(X, row64):
    """X[i] - row for every frame i, the fp64 row applied as its high and low fp32 parts (two passes of the
    rank-1 case of R = M - C V): the result carries the rounding of an fp32 store only."""
    b = X.shape[0]
    ones = torch.ones((b, 1), dtype=torch.float32, device=X.device)
    hi = row64.to(torch.float32)
    lo = (row64 - hi.to(torch.float64)).to(torch.float32)
    R = kernels.project_subtract(X, ones, hi.reshape(1, -1).contiguous())
    return kernels.project_subtract(R, ones, lo.reshape(1, -1).contiguous(), out=R)


class IncrementalModel:
    """State of scikit-learn's ``IncrementalPCA`` (mean, number of samples, singular values, components), updated
    one batch at a time on the device (``sklearn/decomposition/_incremental_pca.py``: ``partial_fit``)."""

    def __init__(self, ncomp):
        self.k = int(ncomp)
        self.n_seen = 0
        self.mean = None            # (p,) fp64
        self.components = None      # (k, p) fp32
        self.svals = None           # (k,) fp64

    def partial_fit(self, X):
        b, p = X.shape
        if self.k > p:
            raise ValueError("n_components=%r invalid for n_features=%d, need more rows than columns for "
                             "IncrementalPCA processing" % (self.k, p))
        if self.n_seen == 0 and self.k > b:
            raise ValueError(f"n_components={self.k} must be less or equal to the batch number of samples {b} for "
                             "the first partial_fit call.")
        col_sum = _column_sums(X)
        n_total = self.n_seen + b
        if self.n_seen == 0:
            col_mean = col_sum / n_total
            stack = _subtract_row_f64(X, col_mean)
        else:
            col_mean = (self.mean * self.n_seen + col_sum) / n_total
            batch_mean = col_sum / b
            Xc = _subtract_row_f64(X, batch_mean)
            corr = np.sqrt((self.n_seen / n_total) * b) * (self.mean - batch_mean)
            basis = (self.components.to(torch.float64) * self.svals[:, None]).to(torch.float32)
            stack = torch.cat((basis, Xc, corr.to(torch.float32).reshape(1, p)), dim=0).contiguous()
        dec = Decomposition(stack, self.k)
        self.components = dec.pcs(self.k)
        self.svals = dec.S[:self.k].clone()
        self.mean = col_mean
        self.n_seen = n_total

    def flipped_components(self):
        """scikit-learn's ``svd_flip(u_based_decision=False)``: the entry of largest magnitude of every component
        is positive."""
        V = self.components
        idx = torch.argmax(V.abs(), dim=1)
        sign = torch.sign(V[torch.arange(V.shape[0], device=V.device), idx])
        sign = torch.where(sign == 0, torch.ones_like(sign), sign)
        return V * sign[:, None]


def pca_incremental(cube, angle_list, batch=0.25, ncomp=1, collapse="median", verbose=True, full_output=False,
                    return_residuals=False, start_time=None, weights=None, **rot_options):
    """Drop-in for ``vip_hci.psfsub.utils_pca.pca_incremental``: ``cube`` / ``angle_list`` are numpy arrays or paths
    of FITS files (``vip_b200.fits``, no astropy needed).  Returns ``frame``; with ``full_output`` ``(frame, model, pcs, medians)`` where ``model`` is the
    :class:`IncrementalModel` (the reference returns its scikit-learn object there); with ``return_residuals``
    the (n, y, x) residual cube."""
    # FITS paths (utils_pca.py:508-519): the data unit of the first HDU as a (big-endian) memory map -- the mini-batches
    # are sliced from it, converted and uploaded one at a time, the file is never read as a whole
    if isinstance(cube, str):
        from ..fits import open_fits
        cube = open_fits(cube, n=0, return_memmap=True, verbose=False)
    if isinstance(angle_list, str):
        from ..fits import open_fits
        angle_list = np.asarray(open_fits(angle_list, verbose=False))
    if not isinstance(cube, np.ndarray):
        raise TypeError("`cube` must be a str (full path on disk) or a numpy array")
    if not isinstance(angle_list, np.ndarray):
        raise TypeError("`angle_list` must be a str (full path on disk) or a numpy array")
    if not cube.ndim > 2:
        raise TypeError("Input array is not a 3d array")
    if cube.ndim != 3:
        raise NotImplementedError("vip_b200.pca_incremental handles 3-d cubes")
    n, y, x = cube.shape
    angle_list = check_pa_vector(angle_list)
    if n != angle_list.shape[0] and not return_residuals:
        raise TypeError("`angle_list` vector has wrong length. It must be the same as the number of frames in the "
                        "cube")
    if not isinstance(ncomp, (int, float)):
        raise TypeError("`ncomp` must be an int or a float in the ADI case")
    if ncomp > n:
        ncomp = min(ncomp, n)
        print("Number of PCs too high (max PCs={}), using {} PCs instead.".format(n, ncomp))
    for key in ("nproc", "interpolation", "ker"):
        rot_options.pop(key, None)
    imlib = rot_options.pop("imlib", "vip-fft")
    _check_rot_options(imlib, rot_options.get("cxy"), rot_options.get("border_mode", "constant"),
                       rot_options.get("edge_blend"), cube.shape)
    mask_val = float(rot_options.get("mask_val", np.nan))
    interp_zeros = bool(rot_options.get("interp_zeros", False))

    dev = _device.require_cuda()
    avail = _device.free_memory_bytes() if isinstance(batch, float) else None
    ranges = batch_ranges(n, batch, cube.nbytes, avail)
    if verbose:
        print("Cube size = {:.3f} GB ({} frames)".format(cube.nbytes / 1e9, n))
        print("Batch size = {} frames ({:.3f} GB), {} batches\n".format(ranges[0][1], cube[:ranges[0][1]].nbytes / 1e9,
                                                                         len(ranges)))

    def upload(i0, i1):
        return _device.to_device_f32(cube[i0:i1], dev).reshape(i1 - i0, y * x)

    model = IncrementalModel(ncomp)
    for i0, i1 in ranges:
        model.partial_fit(upload(i0, i1))
    V = model.flipped_components()

    residuals = np.empty((n, y, x)) if return_residuals else None
    frames = []
    for i0, i1 in ranges:
        Xm = _subtract_row_f64(upload(i0, i1), model.mean)
        Cm = kernels.cross_gram(Xm, V).to(torch.float32).contiguous()
        R = kernels.project_subtract(Xm, Cm, V, out=Xm).reshape(i1 - i0, y, x)
        if return_residuals:
            residuals[i0:i1] = _device.to_host(R)
        else:
            der = derotate_device(R, -angle_list[i0:i1], mask_val=mask_val, interp_zeros=interp_zeros)
            frames.append(collapse_device(der, mode=collapse, w=weights).to(torch.float32))
    if return_residuals:
        return residuals
    medians = torch.stack(frames)
    # np.median over the batch frames: NaN wherever any batch frame is NaN
    frame = collapse_device(medians, "median")
    frame = torch.where(torch.isnan(medians).any(dim=0), torch.full_like(frame, float("nan")), frame)
    frame = _device.to_host(frame).astype(np.float64)
    if full_output:
        pcs = _device.to_host(V).astype(np.float64).reshape(V.shape[0], y, x)
        return frame, model, pcs, _device.to_host(medians).astype(np.float64)
    return frame
