"""Median-ADI / median-RDI subtraction, full-frame mode (``vip_hci/psfsub/medsub.py:91-470``).

A composition of the kernels of the PCA path: temporal median (exact selection kernel) -> broadcast
subtraction -> FFT derotation -> collapse.  Implemented: 3-d cubes, ``mode='fullfr'``, optional
``cube_ref`` with ``collapse_ref`` 'median' / 'mean', ``radius_int``, every ``collapse`` mode.
"""
from dataclasses import dataclass
from enum import Enum
from typing import List, Tuple, Union

import numpy as np
import torch

from .. import kernels
from .._device import to_device_f32, to_host
from ..config.paramenum import ALGO_KEY, Collapse, Imlib, Interpolation
from ..config.utils_param import separate_kwargs_dict
from ..preproc.derotation import derotate_device, _check_rot_options
from ..preproc.parangles import check_pa_vector
from ..preproc.subsampling import collapse_device
from ..var.shapes import circle_mask


@dataclass
class MEDIAN_SUB_Params:
    """Parameters of ``median_sub`` in the reference's declaration order (``medsub.py:61-88``)."""

    cube: np.ndarray = None
    angle_list: np.ndarray = None
    scale_list: np.ndarray = None
    flux_sc_list: np.ndarray = None
    fwhm: float = 4
    radius_int: int = 0
    asize: int = 4
    delta_rot: int = 1
    delta_sep: Union[float, Tuple[float]] = (0.1, 1)
    mode: str = "fullfr"
    nframes: int = 4
    sdi_only: bool = False
    imlib: Enum = Imlib.VIPFFT
    interpolation: Enum = Interpolation.LANCZOS4
    collapse: Enum = Collapse.MEDIAN
    cube_ref: np.ndarray = None
    collapse_ref: str = "median"
    nproc: int = 1
    full_output: bool = False
    verbose: bool = True


def _median_frame_device(cube_dev):
    """``np.median(cube, axis=0)``: the exact selection kernel gives nanmedian; a pixel with any NaN
    sample is NaN for np.median."""
    med = collapse_device(cube_dev, "median")
    return torch.where(torch.isnan(cube_dev).any(dim=0), torch.full_like(med, float("nan")), med)


def _subtract_frame_device(cube_dev, frame_dev):
    """cube[i] - frame for every i, as the rank-1 case of the project-subtract kernel (R = M - C V)."""
    n, H, W = cube_dev.shape
    ones = torch.ones((n, 1), dtype=torch.float32, device=cube_dev.device)
    R = kernels.project_subtract(cube_dev.reshape(n, H * W), ones, frame_dev.reshape(1, H * W).contiguous())
    return R.reshape(n, H, W)


def median_sub_device(cube_dev, angle_list, radius_int=0, collapse="median", cube_ref_dev=None,
                      collapse_ref="median", **rot_options):
    """(cube_out, cube_der, frame) as CUDA tensors for a (n,H,W) fp32 CUDA cube."""
    if cube_ref_dev is None:
        model = _median_frame_device(cube_dev)
    elif "median" in collapse_ref:
        model = _median_frame_device(cube_ref_dev)
    elif "mean" in collapse_ref:
        # np.mean propagates NaNs; the collapse kernel's 'mean' is nanmean
        model = collapse_device(cube_ref_dev, "mean")
        model = torch.where(torch.isnan(cube_ref_dev).any(dim=0), torch.full_like(model, float("nan")), model)
    else:
        raise NotImplementedError("vip_b200.median_sub: collapse_ref must contain 'median' or 'mean' "
                                  "(flux-scaled references are not implemented; no CPU fallback)")
    cube_out = _subtract_frame_device(cube_dev, model)
    cube_der = derotate_device(cube_out, -np.asarray(angle_list, dtype=np.float64),
                               mask_val=rot_options.get("mask_val", np.nan),
                               interp_zeros=rot_options.get("interp_zeros", False))
    if radius_int:
        H, W = cube_dev.shape[1:]
        mask = torch.as_tensor(circle_mask((H, W), radius_int)).to(cube_dev.device)
        cube_out = cube_out.masked_fill(mask[None], 0.0)
        cube_der = cube_der.masked_fill(mask[None], 0.0)
    frame = collapse_device(cube_der, collapse)
    return cube_out, cube_der, frame


def median_sub(*all_args: List, **all_kwargs: dict):
    """Median-ADI / median-RDI: drop-in for ``vip_hci.psfsub.median_sub`` (3-d cubes, ``mode='fullfr'``).

    Returns ``frame`` or, with ``full_output``, ``(cube_out, cube_der, frame)`` (``medsub.py:466-470``)."""
    class_params, rot_options = separate_kwargs_dict(initial_kwargs=all_kwargs, parent_class=MEDIAN_SUB_Params)
    p = rot_options.pop(ALGO_KEY, None)
    if p is None:
        p = MEDIAN_SUB_Params(*all_args, **class_params)
    # by default the masked centre is derotated as zeros (medsub.py:226-229)
    if p.radius_int and len(rot_options) == 0:
        rot_options["mask_val"] = 0
        rot_options["ker"] = 1
        rot_options["interp_zeros"] = True
    if not isinstance(p.cube, np.ndarray) or p.cube.ndim not in (3, 4):
        raise TypeError("Input array is not a 3d or 4d array")
    if p.cube.ndim == 4 or p.scale_list is not None:
        raise NotImplementedError("vip_b200.median_sub: ADI+SDI (4-d cubes) is not implemented on the B200 "
                                  "path yet (no CPU fallback)")
    mode = str(getattr(p.mode, "value", p.mode))
    if mode == "annular":
        raise NotImplementedError("vip_b200.median_sub: mode='annular' is not implemented on the B200 path "
                                  "yet (no CPU fallback)")
    if mode != "fullfr":
        raise RuntimeError("Mode not recognized")
    _check_rot_options(p.imlib, rot_options.get("cxy"), rot_options.get("border_mode", "constant"),
                       rot_options.get("edge_blend"), p.cube.shape)
    angle_list = check_pa_vector(np.asarray(p.angle_list))
    n, y, x = p.cube.shape
    if p.cube_ref is not None and (p.cube_ref.shape[-1] != x or p.cube_ref.shape[-2] != y):
        raise TypeError("Reference cube shape should have same xy dimensions as science cube")
    if n != angle_list.shape[0]:
        raise TypeError("Input vector or parallactic angles has wrong length")
    ref_dev = to_device_f32(p.cube_ref) if p.cube_ref is not None else None
    cube_out, cube_der, frame = median_sub_device(
        to_device_f32(p.cube), angle_list, radius_int=p.radius_int, collapse=p.collapse, cube_ref_dev=ref_dev,
        collapse_ref=p.collapse_ref, **rot_options)
    dt = p.cube.dtype if np.issubdtype(p.cube.dtype, np.floating) else np.float64
    frame = to_host(frame, dtype=dt)
    if p.full_output:
        return to_host(cube_out, dtype=dt), to_host(cube_der, dtype=dt), frame
    return frame
