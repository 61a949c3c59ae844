"""Seeded synthetic ADI cubes shared by tests, smoke() and bench.py.

Stellar halo + quasi-static speckle modes with AR(1) temporal coefficients + read noise +
one injected Gaussian companion that follows the parallactic-angle track.  fp32, positive,
no NaN / exact-zero pixels, non-uniform PA spacing (avoids |dPA| ties in the annular
library selection).
"""
import numpy as np


def _smooth_modes(rng, K, H, W, sigma=2.0):
    """K unit-rms smooth random patterns: white noise low-passed with a Gaussian in Fourier space."""
    fy = np.fft.fftfreq(H)[:, None]
    fx = np.fft.rfftfreq(W)[None, :]
    filt = np.exp(-2.0 * (np.pi * sigma) ** 2 * (fy ** 2 + fx ** 2))
    out = np.empty((K, H, W))
    for j in range(K):
        out[j] = np.fft.irfft2(np.fft.rfft2(rng.standard_normal((H, W))) * filt, s=(H, W))
        out[j] /= out[j].std()
    return out


def pa_track(rng, n, delta_deg, start=10.0):
    d = rng.uniform(0.5, 1.5, n)
    a = np.cumsum(d)
    return (a - a[0]) * (delta_deg / (a[-1] - a[0])) + start


def adi_cube(n, size, ncomp_max=20, delta_deg=90.0, seed=20260101, chunk=64, dtype=np.float32,
             planet_peak=30.0, decay=0.88):
    """Return (cube[n,size,size], angles[n]).  ``decay``: amplitude ratio of consecutive speckle modes (0.88:
    the weakest of 20 modes sits ~4x above the noise floor; closer to 1 keeps more modes above it)."""
    rng = np.random.default_rng(seed)
    H = W = size
    angs = pa_track(rng, n, delta_deg)
    cy = cx = size // 2 if size % 2 == 0 else (size - 1) // 2
    yy, xx = np.mgrid[:H, :W]
    r = np.hypot(yy - cy, xx - cx)
    halo = 1e4 / (1.0 + (r / 4.0) ** 2)
    # exactly ncomp_max time-variable speckle modes, multiplicative on the halo, so that the
    # singular spectrum has a clear gap right after ncomp_max (weakest mode ~4x the noise floor)
    K = ncomp_max
    modes = _smooth_modes(rng, K, H, W) * halo
    # AR(1) temporal coefficients
    phi = 0.9
    ar = np.empty((n, K))
    ar[0] = rng.standard_normal(K)
    for t in range(1, n):
        ar[t] = phi * ar[t - 1] + np.sqrt(1 - phi ** 2) * rng.standard_normal(K)
    coef = 0.05 * (decay ** np.arange(K))[None, :] * (0.5 + ar)
    cube = np.empty((n, H, W), dtype=dtype)
    rp = 0.3 * H
    sig = 4.0 / 2.3548
    modes2 = modes.reshape(K, -1)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        blk = halo[None] + (coef[s:e] @ modes2).reshape(e - s, H, W)
        blk += 3.0 * rng.standard_normal(blk.shape)
        for t in range(s, e):
            th = np.deg2rad(angs[t])
            py, px = cy + rp * np.sin(th), cx + rp * np.cos(th)
            blk[t - s] += planet_peak * np.exp(-((yy - py) ** 2 + (xx - px) ** 2) / (2 * sig ** 2))
        cube[s:e] = blk.astype(dtype)
    return cube, angs
