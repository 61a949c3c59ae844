"""Blob detection on a post-processed frame (drop-in for ``vip_hci.metrics.detection``, modes 'lpeaks' and 'snrmap').

Reference: ``src/vip_hci/metrics/detection.py`` -- ``detection`` :26-382 (``check_blobs`` :126-175).  This is the
acceptance criterion of the reference's own pipeline tests (``tests/helpers.py:38-77``: an injected companion must be
recovered within 3 px).  On the GPU: the S/N map (``snrmap``: one warp per pixel, ``csrc/snr.cu``), the peak mask of the
blob finder (``vb_local_max_mask_f32``, ``csrc/detect.cu``) and the exact-aperture S/N of every candidate.  Host control
logic on a handful of numbers, as in the reference: the sigma-clipped background level of the frame, the ordering /
spacing of at most 20 peaks, and the 2-d Gaussian fit of a (3 fwhm)^2 cut-out per peak (MINPACK through scipy, which is
what astropy's ``LevMarLSQFitter`` calls).  ``mode`` 'log' / 'dog' (scikit-image blob finders), ``matched_filter`` and
plotting are not implemented.
"""
import numpy as np
import torch

from .. import _cabi
from .._device import empty, ptr, require_cuda, stream_ptr
from ..config.utils_conf import sep
from ..var.coords import frame_center
from ..var.shapes import mask_circle
from .snr_source import _frame_to_device, snr, snrmap

GAUSSIAN_FWHM_TO_SIGMA = 1.0 / (2.0 * np.sqrt(2.0 * np.log(2.0)))


def _unsupported(what):
    raise NotImplementedError(f"vip_b200.metrics.detection: {what} is not implemented on the B200 path yet "
                              "(no CPU fallback)")


def sigma_clipped_stats(data, sigma=5.0, maxiters=None):
    """(mean, median, std) after iterative rejection of the points farther than ``sigma`` standard deviations from
    the median (astropy ``sigma_clipped_stats`` defaults: median centre, std, non-finite values masked)."""
    x = np.asarray(data, dtype=np.float64).ravel()
    x = x[np.isfinite(x)]
    it = 0
    while True:
        med, std = np.median(x), np.std(x)
        keep = (x >= med - sigma * std) & (x <= med + sigma * std)
        it += 1
        x = x[keep]
        if keep.all() or (maxiters is not None and it >= maxiters):
            break
    return float(np.mean(x)), float(np.median(x)), float(np.std(x))


def local_max_mask_device(frame_dev, min_distance, threshold):
    """Boolean (H, W) CUDA tensor: pixels that are the maximum of their (2 d + 1)^2 neighbourhood, above
    ``threshold`` and farther than d from the border (``vb_local_max_mask_f32``)."""
    H, W = frame_dev.shape
    mask = empty((H, W), torch.uint8, frame_dev.device)
    _cabi.check(_cabi.lib().vb_local_max_mask_f32(ptr(frame_dev), H, W, int(min_distance), float(threshold),
                                                  ptr(mask), stream_ptr()), "vb_local_max_mask_f32")
    return mask.bool()


def peak_local_max(image, min_distance=1, threshold_abs=None, num_peaks=20):
    """``skimage.feature.peak_local_max(image, min_distance=, threshold_abs=, num_peaks=)``: (npeaks, 2) integer
    (y, x) coordinates by decreasing intensity; the peak mask is computed on the GPU."""
    image = np.asarray(image)
    thr = threshold_abs if threshold_abs is not None else np.nanmin(image)
    dev = require_cuda()
    mask = local_max_mask_device(_frame_to_device(image, dev), min_distance, thr).cpu().numpy()
    yy, xx = np.nonzero(mask)
    order = np.argsort(-image[yy, xx].astype(np.float64), kind="stable")
    out = []
    for y, x in zip(yy[order], xx[order]):                 # ensure_spacing(p_norm=inf), greedy from the brightest
        if all(max(abs(int(y) - a), abs(int(x) - b)) > min_distance for a, b in out):
            out.append((int(y), int(x)))
            if len(out) >= num_peaks:
                break
    return np.array(out, dtype=np.int64).reshape(-1, 2)


def _gauss2d_parts(p, x, y):
    amp, x0, y0, sxd, syd, th = p
    c2, s2 = np.cos(th) ** 2, np.sin(th) ** 2
    s2t, c2t = np.sin(2 * th), np.cos(2 * th)
    xs2, ys2 = sxd ** 2, syd ** 2
    a = 0.5 * (c2 / xs2 + s2 / ys2)
    b = 0.5 * (s2t / xs2 - s2t / ys2)
    cc = 0.5 * (s2 / xs2 + c2 / ys2)
    dx, dy = x - x0, y - y0
    g = amp * np.exp(-(a * dx ** 2 + b * dx * dy + cc * dy ** 2))
    return g, a, b, cc, dx, dy, (c2, s2, s2t, c2t, xs2, ys2)


def fit_gaussian2d(subim, amplitude, x_mean, y_mean, x_stddev, y_stddev, theta=0.0):
    """Levenberg-Marquardt fit of an elliptical 2-d Gaussian (amplitude, x_mean, y_mean, x_stddev, y_stddev, theta) to
    a small image with the analytic derivatives of astropy's ``Gaussian2D`` and its fitter defaults (100 evaluations,
    accuracy 1e-7).  Returns the six fitted parameters."""
    from scipy.optimize import leastsq
    sy, sx = np.indices(subim.shape)
    x, y, z = sx.ravel().astype(np.float64), sy.ravel().astype(np.float64), np.asarray(subim, np.float64).ravel()

    def resid(p):
        return _gauss2d_parts(p, x, y)[0] - z

    def jac(p):
        amp, _, _, sxd, syd, _ = p
        g, a, b, cc, dx, dy, (c2, s2, s2t, c2t, xs2, ys2) = _gauss2d_parts(p, x, y)
        xs3, ys3 = sxd ** 3, syd ** 3
        da_dth = -0.5 * s2t / xs2 + 0.5 * s2t / ys2
        db_dth = c2t / xs2 - c2t / ys2
        quad = lambda da, db, dc: g * -(da * dx ** 2 + db * dx * dy + dc * dy ** 2)      # noqa: E731
        return np.array([g / amp, g * (2.0 * a * dx + b * dy), g * (b * dx + 2.0 * cc * dy),
                         quad(-c2 / xs3, -s2t / xs3, -s2 / xs3), quad(-s2 / ys3, s2t / ys3, -c2 / ys3),
                         quad(da_dth, db_dth, -da_dth)])

    p0 = np.array([amplitude, x_mean, y_mean, x_stddev, y_stddev, theta], dtype=np.float64)
    sol = leastsq(resid, p0, Dfun=jac, col_deriv=True, maxfev=100, ftol=1e-7, xtol=1e-7, gtol=1e-7, full_output=True)
    return tuple(float(v) for v in sol[0])


def detection(array, fwhm=4, psf=None, mode="lpeaks", bkg_sigma=5, matched_filter=False, mask=True, snr_thresh=5,
              nproc=1, plot=True, debug=False, full_output=False, verbose=True, **kwargs):
    """Automatically find point-like sources in a 2-d frame (``metrics/detection.py:26-382``).

    Returns ``(yy, xx)`` of the sources that pass the Gaussian-fit constraints and the S/N threshold, the reference's
    table (a pandas DataFrame with columns y, x, px_snr) with ``full_output``, or ``(0, 0)`` when nothing is found.
    ``plot`` is accepted and ignored (no display on the GPU path)."""
    if array.ndim != 2:
        raise TypeError("Input array is not a frame or 2d array")
    if psf is not None:
        if psf.ndim != 2 and psf.shape[0] < array.shape[0]:
            raise TypeError("Input psf is not a 2d array or has wrong size")
    elif matched_filter:
        raise ValueError("`psf` must be provided when `matched_filter` is True")
    if matched_filter:
        _unsupported("matched_filter=True")
    if fwhm is None:
        if psf is None:
            raise ValueError("`fwhm` or `psf` must be provided")
        cy, cx = frame_center(psf)
        sig0 = 4 * GAUSSIAN_FWHM_TO_SIGMA
        _, _, _, sxd, syd, _ = fit_gaussian2d(psf, psf.max(), cx, cy, sig0, sig0)
        fwhm = float(np.mean([abs(sxd), abs(syd)]) / GAUSSIAN_FWHM_TO_SIGMA)
        if verbose:
            print("FWHM = {:.2f} pxs\n".format(fwhm))
    if mode in ("log", "dog"):
        _unsupported(f"mode={mode!r} (scikit-image blob finders)")
    if mode == "snrmapf":
        _unsupported("mode='snrmapf' (approximated S/N map)")
    if mode not in ("lpeaks", "snrmap"):
        raise ValueError("`mode` not recognized")

    def print_abort():
        if verbose:
            print(sep)
            print("No potential sources found")
            print(sep)

    if mask:
        array = mask_circle(array, radius=fwhm)
    if mode == "lpeaks":
        frame_det = array
        _, median, stddev = sigma_clipped_stats(frame_det, sigma=5, maxiters=None)
        threshold = median + stddev * bkg_sigma
        if debug:
            print("Sigma clipped median = {:.3f}".format(median))
            print("Sigma clipped stddev = {:.3f}".format(stddev))
            print("Background threshold = {:.3f}".format(threshold), "\n")
        pad = 10
        array_fit = np.pad(array, pad, "constant", constant_values=0)
    else:
        frame_det = snrmap(array, fwhm=fwhm, approximated=False, plot=False, nproc=nproc, verbose=verbose)
        threshold = snr_thresh
        pad = 0
        array_fit = array
    coords_temp = peak_local_max(frame_det, threshold_abs=threshold, min_distance=int(np.ceil(fwhm)), num_peaks=20)

    # Gaussian fit of every candidate on a (3 ceil(fwhm) | odd) cut-out (check_blobs, :126-175)
    coords = []
    sig = fwhm * GAUSSIAN_FWHM_TO_SIGMA
    for y, x in coords_temp:
        subsi = 3 * int(np.ceil(fwhm))
        if subsi % 2 == 0:
            subsi += 1
        wing = (subsi - 1) / 2
        scy, scx = y + pad, x + pad
        y0, x0 = int(scy - wing), int(scx - wing)
        y1, x1 = int(scy + wing + 1), int(scx + wing + 1)
        if y0 < 0 or x0 < 0 or y1 > array_fit.shape[0] or x1 > array_fit.shape[1]:
            raise RuntimeError("square cannot be obtained with size={}, y={}, x={}".format(subsi, scy, scx))
        subim = array_fit[y0:y1, x0:x1]
        cy, cx = frame_center(subim)
        amp, xm, ym, sxd, syd, _ = fit_gaussian2d(subim, subim.max(), cx, cy, sig, sig)
        fwhm_y, fwhm_x = syd / GAUSSIAN_FWHM_TO_SIGMA, sxd / GAUSSIAN_FWHM_TO_SIGMA
        mean_fwhm_fit = np.mean([np.abs(fwhm_x), np.abs(fwhm_y)])
        if amp > 0 and np.allclose(xm, cx, atol=2) and np.allclose(ym, cy, atol=2) and \
                np.allclose(mean_fwhm_fit, fwhm, atol=3):
            coords.append((y0 + ym, x0 + xm))
            if debug:
                print("Coordinates (Y,X): {:.3f},{:.3f}".format(y, x))
                print("fit peak = {:.3f}".format(amp))
                print("fwhm_y in px = {:.3f}, fwhm_x in px = {:.3f}".format(fwhm_y, fwhm_x))
                print("mean fit fwhm = {:.3f}".format(mean_fwhm_fit))
    coords = np.array(coords)
    if coords.shape[0] == 0:
        print_abort()
        return 0, 0
    if verbose:
        print("Blobs found:", len(coords))
        print(" ycen   xcen")
        print("------ ------")
        for j in range(len(coords)):
            print("{:.3f} \t {:.3f}".format(coords[j, 0] - pad, coords[j, 1] - pad))
    yy, xx = coords[:, 0] - pad, coords[:, 1] - pad

    yy_final, xx_final, snr_final, snr_list = [], [], [], []
    for y, x in zip(yy, xx):
        if verbose:
            print("")
            print(sep)
            print("X,Y = ({:.1f},{:.1f})".format(x, y))
        snr_value = snr(array, (x, y), fwhm, False, verbose=False)
        snr_list.append(snr_value)
        if snr_value >= snr_thresh:
            yy_final.append(y)
            xx_final.append(x)
            snr_final.append(snr_value)
            if verbose:
                print("S/N = {:.3f}".format(snr_value))
        elif verbose:
            print("S/N constraint NOT fulfilled (S/N = {:.3f})".format(snr_value))
    if verbose:
        print(sep)
    if full_output:
        import pandas as pn
        return pn.DataFrame({"y": yy_final, "x": xx_final, "px_snr": snr_final})
    return np.array(yy_final), np.array(xx_final)
