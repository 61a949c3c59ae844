"""Cube derotation with the FFT three-shear method and the ADI library-selection helpers
(``vip_hci/preproc/derotation.py``).

Host side = geometry and per-frame scalars, computed with the same integer/fp64 expressions as
the reference so that crop windows, the rot90 quadrant and the shear coefficients are bit-exact;
device side = ``vb_derotate_f32`` (``csrc/derotate.cu``).
"""
import numpy as np
import torch

from .. import kernels
from .._device import to_device_f32, to_host


def _center1d(n):
    return n // 2 if n % 2 == 0 else (n - 1) // 2


def rotation_geometry(size):
    """(N, y0) for square frames of ``size`` pixels.

    The reference embeds the frame in a NaN canvas of ``int(1.5*size)`` pixels (parity matched,
    ``derotation.py:150-156``), zero-pads it by 4/1.5 with ``frame_pad`` (rounded, parity kept,
    ``cosmetics.py:206-212``) and, inside ``rotate_fft``, works on an even plane (``derotation.py:
    577-599``).  N is the size of that even working plane and y0 the offset of the frame in it."""
    n15 = int(1.5 * size)
    if n15 % 2 != size % 2:
        n15 += 1
    n4 = int(round(n15 * (4 / 1.5)))
    if n4 % 2 != n15 % 2:
        n4 -= 1
    y0 = int(_center1d(n4) - _center1d(n15)) + int(_center1d(n15) - _center1d(size))
    N = n4 if n4 % 2 == 0 else n4 - 1
    return N, y0


def rotation_scalars(angles):
    """Per-frame (krot, a, b) for rotation angles in degrees (``derotation.py:570-603``): the angle
    is wrapped to [0, 360]; above 45 deg the plane is first rot90'ed ``rint(angle/90)`` times and
    the residual angle lies in [-45, 45]; a = tan(residual/2), b = -sin(residual).

    Vectorised over frames with the reference's own fp64 expressions (same ufuncs, so the same bits as
    the frame-by-frame form ``_rotation_scalars_loop``; tests/test_host_logic.py)."""
    ang = np.array(angles, dtype=np.float64)
    while np.any(ang < 0):                       # the reference wraps with repeated +-360, not with fmod
        ang = np.where(ang < 0, ang + 360, ang)
    while np.any(ang > 360):
        ang = np.where(ang > 360, ang - 360, ang)
    big = ang > 45
    d = ang % 90
    d = np.where(d > 45, -(90 - d), d)
    dangle = np.where(big, d, ang)
    krot = np.where(big, np.rint(ang / 90).astype(np.int64) % 4, 0).astype(np.int32)
    a = np.tan(np.deg2rad(dangle) / 2)
    b = -np.sin(np.deg2rad(dangle))
    return krot, a, b


def _rotation_scalars_loop(angles):
    """Frame-by-frame statement of :func:`rotation_scalars` (the reference's scalar code path)."""
    angles = np.asarray(angles, dtype=np.float64)
    krot = np.zeros(angles.shape[0], dtype=np.int32)
    a = np.zeros(angles.shape[0], dtype=np.float64)
    b = np.zeros(angles.shape[0], dtype=np.float64)
    for i, angle in enumerate(angles):
        while angle < 0:
            angle += 360
        while angle > 360:
            angle -= 360
        if angle > 45:
            dangle = angle % 90
            if dangle > 45:
                dangle = -(90 - dangle)
            krot[i] = int(np.rint(angle / 90)) % 4
        else:
            dangle = angle
        a[i] = np.tan(np.deg2rad(dangle) / 2)
        b[i] = -np.sin(np.deg2rad(dangle))
    return krot, a, b


_VIPFFT_ONLY = ("Only imlib='vip-fft' with edge_blend=None and border_mode='constant' is implemented "
                "on the B200 path (no CPU fallback): got {}")


def derotate_device(cube_dev, rot_angles, mask_val=np.nan, interp_zeros=False, force_direct=False):
    """Rotate frame i of a (n,S,S) fp32 CUDA tensor by ``rot_angles[i]`` degrees."""
    n, S, S2 = cube_dev.shape
    if S != S2:
        raise NotImplementedError("vip_b200 derotation handles square frames only")
    N, y0 = rotation_geometry(S)
    krot, a, b = rotation_scalars(rot_angles)
    mask_val = float(mask_val)
    zero_masked = bool(interp_zeros) and not np.isnan(mask_val)
    return kernels.derotate(cube_dev.contiguous(), krot, a, b, S, N, y0, mask_val=mask_val,
                            zero_masked=zero_masked, force_direct=force_direct)


def _check_rot_options(imlib, cxy, border_mode, edge_blend, shape):
    imlib = str(getattr(imlib, "value", imlib))
    if imlib != "vip-fft":
        raise NotImplementedError(_VIPFFT_ONLY.format(f"imlib={imlib!r}"))
    if edge_blend not in (None, ""):
        raise NotImplementedError(_VIPFFT_ONLY.format(f"edge_blend={edge_blend!r}"))
    if cxy is not None:
        cx, cy = cxy
        # the reference compares against the centre of the padded plane (derotation.py:226-234);
        # any explicit centre therefore only passes when it equals that plane centre
        N, _ = rotation_geometry(shape[-1])
        if (cy, cx) != (N // 2, N // 2):
            raise ValueError("'vip-fft' imlib does not yet allow for custom center to be  provided ")


def cube_derotate(array, angle_list, imlib="vip-fft", interpolation="lanczos4", cxy=None, nproc=1,
                  border_mode="constant", mask_val=np.nan, edge_blend=None, interp_zeros=False, ker=1):
    """Rotate each frame of a cube by ``-angle_list[i]`` (derotation to a common north).

    Drop-in for ``vip_hci.preproc.cube_derotate`` (``derotation.py:331-399``) for the default
    ``imlib='vip-fft'`` path.  ``nproc`` is accepted and ignored (frames are processed in parallel
    on the GPU); the output keeps the input dtype."""
    if array.ndim != 3:
        raise TypeError("Input array is not a cube or 3d array.")
    _check_rot_options(imlib, cxy, border_mode, edge_blend, array.shape)
    angles = -np.asarray(angle_list, dtype=np.float64)
    on_device = isinstance(array, torch.Tensor)
    dev = array.float() if on_device else to_device_f32(array)
    out = derotate_device(dev, angles, mask_val=mask_val, interp_zeros=interp_zeros)
    if on_device:
        return out
    return to_host(out, dtype=array.dtype)


def frame_rotate(array, angle, imlib="vip-fft", interpolation="lanczos4", cxy=None, border_mode="constant",
                 mask_val=np.nan, edge_blend=None, interp_zeros=False, ker=1):
    """Rotate one frame by ``angle`` degrees (``derotation.py:51-328``, vip-fft branch); float64 out."""
    if array.ndim != 2:
        raise TypeError("Input array is not a frame or 2d array")
    _check_rot_options(imlib, cxy, border_mode, edge_blend, array.shape)
    dev = to_device_f32(array[None])
    out = derotate_device(dev, np.array([angle], dtype=np.float64), mask_val=mask_val,
                          interp_zeros=interp_zeros)
    return to_host(out[0], dtype=np.float64)


# ------------------------------------------------------------------------------------------------
# library selection for annular ADI (integer results: kept on the host in numpy, bit-exact)
# ------------------------------------------------------------------------------------------------


def _find_indices_adi(angle_list, frame, thr, nframes=None, out_closest=False, truncate=False,
                      max_frames=200):
    """Indices of the frames kept in the library of ``frame`` given a PA threshold ``thr``
    (``derotation.py:410-496``)."""
    n = angle_list.shape[0]
    pa_f = angle_list[frame]
    # first earlier frame closer than thr (else `frame`); first later frame farther than thr (else n)
    index_prev = frame
    for i in range(frame):
        if np.abs(pa_f - angle_list[i]) < thr:
            index_prev = i
            break
    index_foll = n
    for k in range(frame, n):
        if np.abs(angle_list[k] - pa_f) > thr:
            index_foll = k
            break
    if out_closest:
        return index_prev, index_foll - 1
    if nframes is not None:
        window = nframes // 2
        lo = max(index_prev - window, 0)
        hi = min(index_foll + window, n)
        return np.array(list(range(lo, index_prev)) + list(range(index_foll, hi)), dtype="int32")
    kept = list(range(0, index_prev)) + list(range(index_foll, n))
    indices = np.array(kept, dtype="int32")
    if truncate:
        limit = min(n - 1, max_frames)
        everything = np.array(kept)
        if len(everything) > limit:
            d_pa = np.abs(angle_list[everything] - pa_f)
            indices = np.sort(everything[np.argsort(d_pa)][:limit])
    return indices


def _compute_pa_thresh(ann_center, fwhm, delta_rot=1):
    """PA threshold in degrees (``derotation.py:499-504``)."""
    return np.rad2deg(2 * np.arctan(delta_rot * fwhm / (2 * ann_center)))


def _define_annuli(angle_list, ann, n_annuli, fwhm, radius_int, annulus_width, delta_rot, n_segments,
                   verbose, strict=False):
    """(pa_threshold, inner_radius, ann_center) of annulus ``ann`` (``derotation.py:507-539``)."""
    if ann == n_annuli - 1:
        inner_radius = radius_int + (ann * annulus_width - 1)
    else:
        inner_radius = radius_int + ann * annulus_width
    ann_center = inner_radius + (annulus_width / 2)
    pa_threshold = _compute_pa_thresh(ann_center, fwhm, delta_rot)
    mid_range = np.abs(np.amax(angle_list) - np.amin(angle_list)) / 2
    if pa_threshold >= mid_range - mid_range * 0.1:
        new_pa_th = float(mid_range - mid_range * 0.1)
        if strict:
            if int(verbose) > 1:
                print("WARNING: PA threshold {:.2f} is too big, recommended  value for annulus {:.0f}: "
                      "{:.2f}".format(pa_threshold, ann, new_pa_th))
        else:
            print("PA threshold {:.2f} is likely too big, will be set to {:.2f}".format(pa_threshold,
                                                                                      new_pa_th))
            pa_threshold = new_pa_th
    if int(verbose):
        if pa_threshold > 0:
            print("Ann {}    PA thresh: {:5.2f}    Ann center: {:3.0f}    N segments: {} ".format(
                ann + 1, pa_threshold, ann_center, n_segments))
        else:
            print("Ann {}    Ann center: {:3.0f}    N segments: {} ".format(ann + 1, ann_center,
                                                                          n_segments))
    return pa_threshold, inner_radius, ann_center
