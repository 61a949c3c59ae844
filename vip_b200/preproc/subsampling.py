"""Temporal / spectral collapse of a cube (``vip_hci/preproc/subsampling.py:30-116``)."""
import numpy as np
import torch

from .. import kernels
from .._device import to_device_f32, to_host

_MODES = ("median", "mean", "sum", "max", "trimmean", "absmean", "wmean")


def _trim_window(N, n):
    """(k, n) of ``np.sort(...)[k:k+n]`` used by 'trimmean' (``subsampling.py:86-90``)."""
    k = (N - n) // 2
    if N % 2 != n % 2:
        n += 1
    return k, n


def collapse_device(cube_dev, mode="median", n=50, w=None):
    """(N, H, W) fp32 CUDA tensor -> (H, W) CUDA tensor (fp64 for 'wmean')."""
    mode = str(getattr(mode, "value", mode))
    N, H, W = cube_dev.shape
    flat = cube_dev.reshape(N, H * W)
    if mode == "wmean":
        if w is None:
            raise ValueError("Weights have to be provided for weighted mean mode")
        if len(w) != N:
            raise TypeError("Weights need same length as cube")
        out = kernels.collapse(flat, "wmean", w=np.asarray(w, dtype=np.float64))
    elif mode == "trimmean":
        k, nn = _trim_window(N, n)
        out = kernels.collapse(flat, "trimmean", trim_k=k, trim_n=nn)
    elif mode in _MODES:
        out = kernels.collapse(flat, mode)
    else:
        raise TypeError("mode not recognized")
    return out.reshape(H, W)


def cube_collapse(cube, mode="median", n=50, w=None):
    """Collapse a 3-d cube (axis 0) or 4-d cube (axis 1) with NaN-aware statistics.

    Same signature and semantics as the reference; ``cube`` is a numpy array (or a CUDA tensor,
    in which case a CUDA tensor is returned)."""
    on_device = isinstance(cube, torch.Tensor)
    if cube.ndim not in (3, 4):
        raise TypeError("The input array is not a cube or 3d array.")
    if on_device:
        dev = cube.float()
    else:
        dev = to_device_f32(cube)
    if cube.ndim == 3:
        out = collapse_device(dev, mode, n, w)
    else:
        out = torch.stack([collapse_device(dev[j], mode, n, w) for j in range(dev.shape[0])])
    if on_device:
        return out
    res = to_host(out)
    mode = str(getattr(mode, "value", mode))
    if mode != "wmean" and cube.dtype != np.float32 and np.issubdtype(cube.dtype, np.floating):
        res = res.astype(cube.dtype)
    return res
