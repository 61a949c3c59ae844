"""S/N of a test resolution element and S/N map on the B200 (drop-ins for ``vip_hci.metrics.snr`` / ``snrmap``).

Reference: ``src/vip_hci/metrics/snr_source.py`` -- ``snrmap`` :32-204, ``indep_ap_centers`` :229-318,
``snr`` :321-455.  The aperture fluxes (photutils ``aperture_photometry(method='exact')`` in the reference) and, for
the map, the whole per-pixel S/N evaluation run in CUDA kernels (``csrc/snr.cu``); the aperture geometry of a
single ``snr`` call is host logic with the reference's own expressions.  No CPU fallback.
"""
import numpy as np
import torch

from .. import _cabi
from .._device import empty, ptr, require_cuda, stream_ptr
from ..config.utils_conf import check_array
from ..var.coords import frame_center


def _unsupported(what):
    raise NotImplementedError(f"vip_b200.metrics: {what} is not implemented on the B200 path yet (no CPU fallback)")


def indep_ap_centers(array, source_xy, fwhm, exclude_negative_lobes=False, exclude_theta_range=None, no_gap=False):
    """Centres (yy, xx) of the independent apertures at the separation of ``source_xy`` (``snr_source.py:229-318``)."""
    sourcex, sourcey = source_xy
    centery, centerx = frame_center(array)
    sep = np.sqrt((centery - float(sourcey)) ** 2 + (centerx - float(sourcex)) ** 2)
    theta_0 = np.rad2deg(np.arctan2(sourcey - centery, sourcex - centerx))
    if exclude_theta_range is not None:
        exc_theta_range = list(exclude_theta_range)
    if not sep > (fwhm / 2):
        raise RuntimeError("`source_xy` is too close to the frame center")
    sign = -1                                   # clockwise
    if exclude_theta_range is not None:
        if theta_0 > exc_theta_range[0] and theta_0 < exc_theta_range[1]:
            exc_theta_range[0] += 360
        while theta_0 < exc_theta_range[1]:
            theta_0 += 360
    theta = theta_0
    angle = np.arcsin(fwhm / 2.0 / sep) * 2
    number_apertures = int(np.floor(2 * np.pi / angle))
    if no_gap:
        number_apertures += 1
    yy, xx = [sourcey - centery], [sourcex - centerx]
    yy_all, xx_all = np.zeros(number_apertures), np.zeros(number_apertures)
    cosangle, sinangle = np.cos(angle), np.sin(angle)
    xx_all[0], yy_all[0] = sourcex - centerx, sourcey - centery
    for i in range(number_apertures - 1):
        xx_all[i + 1] = cosangle * xx_all[i] - sign * sinangle * yy_all[i]
        yy_all[i + 1] = cosangle * yy_all[i] + sign * sinangle * xx_all[i]
        theta += sign * np.rad2deg(angle)
        if exclude_negative_lobes and (i == 0 or i == number_apertures - 2):
            continue
        if exclude_theta_range is None or theta < exc_theta_range[0] or theta > exc_theta_range[1]:
            xx.append(cosangle * xx_all[i] - sign * sinangle * yy_all[i])
            yy.append(cosangle * yy_all[i] + sign * sinangle * xx_all[i])
    return np.array(yy) + centery, np.array(xx) + centerx


def _frame_to_device(array, dev):
    if isinstance(array, torch.Tensor):
        return array.to(device=dev, dtype=torch.float32).contiguous()
    return torch.from_numpy(np.ascontiguousarray(array, dtype=np.float32)).to(dev)


def aperture_sums_device(frame_dev, xs, ys, r):
    """Exact circular-aperture sums of a (H,W) fp32 CUDA frame at centres (xs, ys): fp64 CUDA tensor."""
    lib = _cabi.lib()
    H, W = frame_dev.shape
    dev = frame_dev.device
    d_x = torch.as_tensor(np.asarray(xs, dtype=np.float64)).to(dev)
    d_y = torch.as_tensor(np.asarray(ys, dtype=np.float64)).to(dev)
    out = empty((d_x.numel(),), torch.float64, dev)
    _cabi.check(lib.vb_aperture_sums_f64(ptr(frame_dev), H, W, ptr(d_x), ptr(d_y), d_x.numel(), float(r), ptr(out),
                                         stream_ptr()), "vb_aperture_sums_f64")
    return out


def snr_points_device(frame_dev, xs, ys, fwhm, frame2_dev=None, use2alone=False, exclude_negative_lobes=False):
    """S/N and source flux at integer pixel positions of a CUDA frame (one warp per position): two fp64 tensors."""
    lib = _cabi.lib()
    H, W = frame_dev.shape
    dev = frame_dev.device
    d_x = torch.as_tensor(np.asarray(xs, dtype=np.int32)).to(dev)
    d_y = torch.as_tensor(np.asarray(ys, dtype=np.int32)).to(dev)
    npts = d_x.numel()
    out = empty((npts,), torch.float64, dev)
    flux = empty((npts,), torch.float64, dev)
    cy, cx = frame_center((H, W))
    _cabi.check(lib.vb_snr_points_f64(ptr(frame_dev), ptr(frame2_dev), H, W, ptr(d_x), ptr(d_y), npts, float(fwhm),
                                      float(cy), float(cx), int(bool(exclude_negative_lobes)), int(bool(use2alone)),
                                      ptr(out), ptr(flux), stream_ptr()), "vb_snr_points_f64")
    return out, flux


def snr(array, source_xy, fwhm, full_output=False, array2=None, use2alone=False, exclude_negative_lobes=False,
        exclude_theta_range=None, plot=False, verbose=False):
    """S/N of a test resolution element (``snr_source.py:321-455``).  Returns ``snr`` or, with ``full_output``,
    ``(sourcey, sourcex, f_source, fluxes, snr)``."""
    check_array(array, dim=2, msg="array")
    if not isinstance(source_xy, tuple):
        raise TypeError("`source_xy` must be a tuple of floats")
    if array2 is not None and not array2.shape == array.shape:
        raise TypeError("`array2` has not the same shape as input array")
    if plot:
        _unsupported("plot=True")
    sourcex, sourcey = source_xy
    yy, xx = indep_ap_centers(array, source_xy, fwhm, exclude_negative_lobes, exclude_theta_range)
    rad = fwhm / 2.0
    dev = require_cuda()
    fluxes = aperture_sums_device(_frame_to_device(array, dev), xx, yy, rad).cpu().numpy()
    if array2 is not None:
        fluxes2 = aperture_sums_device(_frame_to_device(array2, dev), xx, yy, rad).cpu().numpy()
        fluxes = np.concatenate(([fluxes[0]], fluxes2[:])) if use2alone else np.concatenate((fluxes, fluxes2))
    f_source = fluxes[0].copy()
    fluxes = fluxes[1:]
    n2 = fluxes.shape[0]
    backgr_apertures_std = fluxes.std(ddof=1)
    snr_vale = (f_source - fluxes.mean()) / (backgr_apertures_std * np.sqrt(1 + (1 / n2)))
    if verbose:
        print("S/N for the given pixel = {:.3f}".format(snr_vale))
        print("Integrated flux in FWHM test aperture = {:.3f}".format(f_source))
        print("Mean of background apertures integrated fluxes = {:.3f}".format(fluxes.mean()))
        print("Std-dev of background apertures integrated fluxes = {:.3f}".format(backgr_apertures_std))
    if full_output:
        return sourcey, sourcex, f_source, fluxes, snr_vale
    return snr_vale


def snrmap(array, fwhm, approximated=False, plot=False, known_sources=None, nproc=None, array2=None,
           use2alone=False, exclude_negative_lobes=False, verbose=True, **kwargs):
    """S/N map (``snr_source.py:32-204``): ``snr`` at every non-zero pixel of the annulus
    fwhm <= r < fwhm + min(H, W) / 2 - 1.5 fwhm, all pixels in ONE kernel launch (``nproc`` is accepted and ignored).
    ``approximated`` and ``known_sources`` are not implemented."""
    check_array(array, dim=2, msg="array")
    if approximated:
        _unsupported("approximated=True")
    if known_sources is not None:
        _unsupported("`known_sources`")
    if plot:
        _unsupported("plot=True")
    if array2 is not None and not array2.shape == array.shape:
        raise TypeError("`array2` has not the same shape as input array")
    sizey, sizex = array.shape
    snrmap_array = np.zeros_like(array)
    width = min(sizey, sizex) / 2 - 1.5 * fwhm
    cy, cx = frame_center(array)
    yg, xg = np.mgrid[:sizey, :sizex]
    rad = np.sqrt((xg - cx) ** 2 + (yg - cy) ** 2)
    # the reference masks array * annulus and turns it into a boolean mask: zero-valued pixels drop out as well
    mask = np.asarray(np.asarray(array) * ((rad >= fwhm) & (rad < fwhm + width))).astype(bool)
    yy, xx = np.where(mask)
    if yy.size == 0:
        return snrmap_array
    dev = require_cuda()
    f2 = _frame_to_device(array2, dev) if array2 is not None else None
    values, _ = snr_points_device(_frame_to_device(array, dev), xx, yy, fwhm, f2, use2alone, exclude_negative_lobes)
    snrmap_array[yy, xx] = values.cpu().numpy()
    if verbose:
        print("S/N map created on the GPU (vip_b200)")
    return snrmap_array
