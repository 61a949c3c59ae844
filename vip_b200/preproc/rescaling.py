"""SDI wavelength (de)scaling with the FFT method, as separable linear operators on the GPU.

Reference: ``src/vip_hci/preproc/rescaling.py`` -- ``cube_rescaling_wavelengths`` :324-475,
``frame_rescaling`` (vip-fft branch) :641-673, ``scale_fft`` :1114-1217, ``check_scal_vector`` :767-794, ``_find_indices_sdi`` :916-988.

``scale_fft`` is a fixed linear map per (frame size, scale): zero-pad to N+2kd, FFT, crop/pad the
spectrum to N+2kf, inverse FFT, crop/embed back to N -- the same 1-d complex operator L along both
axes, so  out = Re(L X L^T) = Lr X Lr^T - Li X Li^T  (SURVEY.md 8a-V3).  L is built once per distinct
scale on the host by pushing the identity through the reference's own 1-d index logic (so kd, kf,
crop offsets and the odd-size embedding are inherited exactly); the two-sided products run on the
GPU as batched GEMMs (``vb_gemm_f32``).  The reference feeds a float32 canvas into its complex128
FFTs, so fp32 GEMMs match it at the 1e-6 level.
"""
from functools import lru_cache

import numpy as np


def check_scal_vector(scal_vec):
    """Normalise scaling factors so that the smallest is 1 (``rescaling.py:767-794``)."""
    if not isinstance(scal_vec, (list, np.ndarray)):
        raise TypeError("`Scal_vec` is neither a list or an np.ndarray")
    scal_vec = np.array(scal_vec)
    if scal_vec.min() != 1:
        scal_vec = scal_vec / scal_vec.min()
    return scal_vec


def _find_indices_sdi(scal, dist, index_ref, fwhm, delta_sep=1, nframes=None, debug=False):
    """Spectral channels usable as PSF-model library for channel ``index_ref`` at separation ``dist`` px: those whose
    radial motion after rescaling is at least ``delta_sep`` mean FWHM, on either side of the reference wavelength
    (``rescaling.py:916-988``).  Integer index logic, kept on the host in fp64 with the reference's expressions so
    that the lists are bit-exact; ``nframes`` keeps a window of adjacent channels around the reference one."""
    scal = np.asarray(scal)
    ref = scal[index_ref]
    # motion (in FWHM) of a companion at dist +/- one resolution element, for shorter / longer wavelengths
    shorter = (ref - scal) / ref * ((dist + fwhm * delta_sep) / fwhm) >= delta_sep
    longer = (scal - ref) / ref * ((dist - fwhm * delta_sep) / fwhm) >= delta_sep
    indices = np.nonzero(shorter | longer)[0]
    if debug:
        print("dist: {}, index_ref: {}, indices: {}".format(dist, index_ref, indices))
    if indices.size == 0:
        raise RuntimeError("No frames left after radial motion threshold. Try decreasing the value of `delta_sep`")
    if nframes is not None:
        n_short = int(shorter.sum())
        half = nframes // 2
        if n_short - half < 0 or n_short + half > indices[-1]:
            half = nframes
        indices = indices[max(0, n_short - half):min(scal.size, n_short + half)]
        if indices.size < 2:
            raise RuntimeError("No frames left after radial motion threshold. Try decreasing the value of "
                               "`delta_sep` or `nframes`")
    return indices


def _scale_fft_1d_operator(dim, scale):
    """Complex (dim, dim) matrix L of the 1-d pipeline inside ``scale_fft(..., ori_dim=True)`` for an
    even ``dim`` (``rescaling.py:1137-1215``)."""
    if scale == 1:
        return np.eye(dim, dtype=np.complex128)
    kd_array = np.arange(dim / 2 + 1, dtype=int)
    yy = dim / 2 * (scale - 1) + kd_array.astype(float) * scale
    kf_array = np.round(yy).astype(int)
    imin = np.nanargmin(np.abs(yy - kf_array))
    kd, kf = int(kd_array[imin]), int(kf_array[imin])
    dim_p, dim_pp = dim + 2 * kd, dim + 2 * kf
    canvas = np.zeros((dim_p, dim), dtype=np.complex128)
    canvas[kd:kd + dim] = np.eye(dim)
    spec = np.fft.fftshift(np.fft.fft(canvas, axis=0), axes=0)
    if dim_pp > dim_p:
        big = np.zeros((dim_pp, dim), dtype=np.complex128)
        big[(dim_pp - dim_p) // 2:(dim_pp + dim_p) // 2] = spec
        spec = big
    else:
        spec = spec[kd - kf:kd - kf + dim_pp]
    res = np.fft.ifft(np.fft.fftshift(spec, axes=0), axis=0)          # (dim_pp, dim)
    if dim_pp > dim:
        return res[kf:kf + dim]
    out = np.zeros((dim, dim), dtype=np.complex128)
    out[-kf:-kf + dim_pp] = res
    return out


@lru_cache(maxsize=512)
def rescale_operator(size, scale):
    """Complex (size, size) operator L of ``frame_rescaling(frame, scale=scale, imlib='vip-fft')`` about
    the frame centre for a square frame of ``size`` pixels:  out = Re(L X L^T).  Odd sizes are embedded
    at [1:, 1:] of an even canvas and cropped back, as the reference does (``rescaling.py:648-673``)."""
    scale = float(scale)
    if size % 2 == 0:
        return _scale_fft_1d_operator(size, scale)
    L = _scale_fft_1d_operator(size + 1, scale)
    return L[1:, 1:]


def padded_size(size, max_scale):
    """Frame size after the reflect padding of ``cube_rescaling_wavelengths`` (``rescaling.py:433-443``)."""
    if max_scale <= 1:
        return size
    new = int(np.ceil(max_scale * size))
    if (new - size) % 2 != 0:
        new += 1
    return new


def crop_window(big, size, center):
    """(y0, y1) of ``get_square(frame(big x big), size, center, center)`` (``var/shapes.py:255-350``)."""
    if big % 2 != size % 2:
        size += 1
    wing = (size - 1) / 2
    return int(center - wing), int(center + wing + 1)
