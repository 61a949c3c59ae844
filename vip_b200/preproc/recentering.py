"""Sub-pixel frame shifts with the Fourier phase-ramp method (``vip_hci/preproc/recentering.py``:
``frame_shift`` :30-255 -- the ``imlib='vip-fft'`` branch :122-189 -- and ``cube_shift`` :257-305).

Host side = the reference's integer geometry (padding by ``ceil(max|shift|)``, squaring, evening);
device side = per-frame Toeplitz operators (``vb_shift_operators_f32``), two batched GEMMs
(``vb_gemm_f32``) and the Nyquist checkerboard term (``vb_checker_correct_f32``), see ``csrc/shift.cu``.
"""
import numpy as np
import torch

from .. import kernels
from .._device import to_device_f32, to_host

_VIPFFT_ONLY = ("Only imlib='vip-fft' is implemented on the B200 path (no CPU fallback): got imlib={!r}")
_OPERATOR_BYTES_MAX = 1 << 30


def shift_plane_size(ny, nx, shift_y, shift_x):
    """Size of the even square plane the reference shifts in (``recentering.py:125-169``)."""
    npad = int(np.ceil(np.amax(np.abs([shift_y, shift_x]))))
    npix = max(int(ny + 2 * npad), int(nx + 2 * npad))
    return npix + (npix % 2)


def cube_shift_device(cube_dev, shift_y, shift_x):
    """Shift frame i of a (n,ny,nx) fp32 CUDA tensor by (shift_y[i], shift_x[i]) pixels."""
    n, ny, nx = cube_dev.shape
    shift_y = np.asarray(shift_y, dtype=np.float64)
    shift_x = np.asarray(shift_x, dtype=np.float64)
    if shift_y.shape != (n,) or shift_x.shape != (n,):
        raise TypeError("shift_y / shift_x need one value per frame")
    nplane = np.array([shift_plane_size(ny, nx, sy, sx) for sy, sx in zip(shift_y, shift_x)], dtype=np.int32)
    coef = np.sin(np.pi * shift_x) * np.sin(np.pi * shift_y) / nplane.astype(np.float64) ** 2
    cube_dev = cube_dev.contiguous()
    out = torch.empty_like(cube_dev)
    per_frame = 4 * (ny * ny + nx * nx + ny * nx)
    chunk = max(1, min(n, _OPERATOR_BYTES_MAX // per_frame))
    for f0 in range(0, n, chunk):
        f1 = min(n, f0 + chunk)
        X = cube_dev[f0:f1]
        Tx = kernels.shift_operators(shift_x[f0:f1], nplane[f0:f1], nx, X.device)     # (c,nx,nx)
        Ty = kernels.shift_operators(shift_y[f0:f1], nplane[f0:f1], ny, X.device)     # (c,ny,ny)
        tmp = torch.empty_like(X)
        kernels.gemm(X, Tx, tmp, trans_b=True)             # rows:    tmp = X Tx^T
        kernels.gemm(Ty, tmp, out[f0:f1])                  # columns: out = Ty tmp
        kernels.checker_correct(X, out[f0:f1], coef[f0:f1])
    return out


def _check_imlib(imlib):
    imlib = str(getattr(imlib, "value", imlib))
    if imlib != "vip-fft":
        if imlib in ("ndimage-fourier", "ndimage-interp", "opencv"):
            raise NotImplementedError(_VIPFFT_ONLY.format(imlib))
        raise ValueError("Image transformation library not recognized")


def frame_shift(array, shift_y, shift_x, imlib="vip-fft", interpolation="lanczos4", border_mode="reflect"):
    """Shift a 2-d array by (shift_y, shift_x) pixels: drop-in for ``vip_hci.preproc.frame_shift`` with
    the default ``imlib='vip-fft'`` (zero ``border_mode``, as in the reference for this imlib)."""
    if not isinstance(array, np.ndarray) or array.ndim != 2:
        raise TypeError("Input array is not a frame or 2d array")
    _check_imlib(imlib)
    out = cube_shift_device(to_device_f32(array[None]), [shift_y], [shift_x])[0]
    # the reference multiplies the spectrum by a complex128 ramp: float64 out whatever the input dtype
    return to_host(out, dtype=np.float64)


def cube_shift(cube, shift_y, shift_x, imlib="vip-fft", interpolation="lanczos4", border_mode="reflect",
               nproc=None):
    """Shift every frame of a cube: drop-in for ``vip_hci.preproc.cube_shift`` (``recentering.py:257-305``).
    ``shift_y`` / ``shift_x``: one value per frame, or scalars applied to all frames.  ``nproc`` is accepted
    and ignored (frames are processed in parallel on the GPU); the output keeps the input dtype."""
    on_device = isinstance(cube, torch.Tensor)
    if cube.ndim != 3:
        raise TypeError("Input array is not a cube or 3d array")
    _check_imlib(imlib)
    n = cube.shape[0]
    if np.isscalar(shift_x):
        shift_x = np.ones([n]) * shift_x
    if np.isscalar(shift_y):
        shift_y = np.ones([n]) * shift_y
    dev = cube.float() if on_device else to_device_f32(cube)
    out = cube_shift_device(dev, shift_y, shift_x)
    if on_device:
        return out
    return to_host(out, dtype=cube.dtype)
