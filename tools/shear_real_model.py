"""fp64 numpy model of the REAL-plane three-shear formulation used by csrc/derotate.cu (packed kernels).

Each 1-D shear  y = IFFT(FFT(x) * exp(-2 pi i s f))  of a REAL line x gives  Re(y) + i * beta * (-1)^n
with beta = X[N/2] sin(pi s) / N  (only the Nyquist bin, f = -1/2, breaks Hermitian symmetry).  The
intermediate planes are therefore stored as real planes plus per-line Nyquist scalars, and two real
lines share one complex transform (z = x_a + i x_b).  This script proves the bookkeeping against the
oracle (complex planes, reference semantics) to ~1e-13.

    python tools/shear_real_model.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import vip_oracle  # noqa: E402
from vip_b200.preproc.derotation import rotation_geometry, rotation_scalars  # noqa: E402


def signed_freq(N):
    k = np.arange(N)
    return np.where(k < N // 2, k, k - N) / N          # Nyquist -> -1/2 (numpy fftfreq)


def pair_shear(za, zb, sa, sb):
    """Two real lines through ONE complex FFT; returns Re parts and the two Nyquist scalars."""
    N = za.shape[0]
    f = signed_freq(N)
    Z = np.fft.fft(za + 1j * zb)
    Zc = np.conj(Z[(-np.arange(N)) % N])
    pa = np.exp(-2j * np.pi * sa * f)
    pb = np.exp(-2j * np.pi * sb * f)
    nyq_a = Z[N // 2].real * pa[N // 2].imag / N       # beta of line a
    nyq_b = Z[N // 2].imag * pb[N // 2].imag / N
    pa[N // 2] = pa[N // 2].real
    pb[N // 2] = pb[N // 2].real
    W = Z * (pa + pb) / 2 + Zc * (pa - pb) / 2
    w = np.fft.ifft(W)
    return w.real, w.imag, nyq_a, nyq_b


def rotate_real_model(frame, angle):
    S = frame.shape[0]
    N, y0 = rotation_geometry(S)
    krot, a, b = rotation_scalars(np.array([angle]))
    k, a, b = int(krot[0]), float(a[0]), float(b[0])
    # plane (N+1)^2 with the frame at [y0, y0+S)^2, rot90 k times, last row/col dropped
    P = np.zeros((N + 1, N + 1))
    P[y0:y0 + S, y0:y0 + S] = frame
    P = np.rot90(P, k)[:-1, :-1]
    # primed coordinates: n' = n - y0 (mod N) for rows and columns
    idx = (np.arange(N) + y0) % N
    alt = (-1.0) ** np.arange(N)
    # pass 1: rows r' in [0, S], shift a*(i - N/2)
    A = np.zeros((S + 1, N))
    beta = np.zeros(S + 2)
    rows = list(range(S + 1)) + [None]
    for r in range(0, S + 1, 2):
        xa = P[y0 + r][idx]
        xb = P[y0 + r + 1][idx] if r + 1 <= S else np.zeros(N)
        sa = a * (y0 + r - N / 2)
        sb = a * (y0 + r + 1 - N / 2)
        ya, yb, na, nb = pair_shear(xa, xb, sa, sb)
        A[r] = ya
        beta[r] = na
        if r + 1 <= S:
            A[r + 1] = yb
            beta[r + 1] = nb
    beta = beta[:S + 1]
    sigma_beta = np.sum(alt[:S + 1] * beta)
    # aux: Cs[r'] = Re sum_n shear_{s_n}(beta)[r'],  multiplier M[k] = sum_u exp(-2 pi i b u f_k)
    f = signed_freq(N)
    u = np.arange(N) - N / 2
    theta = 2 * np.pi * b * f
    with np.errstate(invalid="ignore", divide="ignore"):
        M = np.exp(1j * theta / 2) * np.sin(theta * N / 2) / np.sin(theta / 2)
    M[0] = N
    M_chk = np.exp(-1j * np.outer(theta, u)).sum(axis=1)
    assert np.allclose(M, M_chk, atol=1e-8)
    bcol = np.zeros(N)
    bcol[:S + 1] = beta
    Cs = np.fft.ifft(np.fft.fft(bcol) * M).real[:S]
    # pass 2: columns n' in [0, N), input rows [0, S], output rows [0, S)
    Pm = np.zeros((S, N))
    gamma = np.zeros(N)
    for n in range(0, N, 2):
        ca = np.zeros(N)
        cb = np.zeros(N)
        ca[:S + 1] = A[:, n]
        cb[:S + 1] = A[:, n + 1]
        sa = b * (((n + y0) % N) - N / 2)
        sb = b * (((n + 1 + y0) % N) - N / 2)
        ya, yb, na, nb = pair_shear(ca, cb, sa, sb)
        da = sigma_beta * np.sin(np.pi * sa) / N
        db = sigma_beta * np.sin(np.pi * sb) / N
        Pm[:, n] = ya[:S] - alt[n] * alt[:S] * da
        Pm[:, n + 1] = yb[:S] - alt[n + 1] * alt[:S] * db
        gamma[n] = na
        gamma[n + 1] = nb
    Gam = np.sum(alt * gamma)
    # pass 3: rows r' in [0, S), outputs n' in [0, S)
    out = np.zeros((S, S))
    for r in range(0, S, 2):
        sa = a * (y0 + r - N / 2)
        sb = a * (y0 + r + 1 - N / 2)
        ya, yb, _, _ = pair_shear(Pm[r], Pm[r + 1], sa, sb)
        ca = (alt[r] * Gam + Cs[r]) * np.sin(np.pi * sa) / N
        cb = (alt[r + 1] * Gam + Cs[r + 1]) * np.sin(np.pi * sb) / N
        out[r] = ya[:S] - alt[:S] * ca
        out[r + 1] = yb[:S] - alt[:S] * cb
    return out


if __name__ == "__main__":
    rng = np.random.default_rng(3)
    worst = 0.0
    for S in (32, 64):
        for angle in (7.3, -33.1, 58.0, 135.0, 200.5, 315.0, 44.9):
            # checkerboard-heavy content makes the Nyquist terms large
            fr = rng.normal(size=(S, S)) + 5.0 * ((np.add.outer(np.arange(S), np.arange(S)) % 2) * 2 - 1)
            ref = vip_oracle.frame_rotate(fr, angle)
            got = rotate_real_model(fr, angle)
            err = np.max(np.abs(got - ref)) / np.max(np.abs(ref))
            worst = max(worst, err)
            print(f"S={S} angle={angle:7.2f}  rel err {err:.2e}")
    assert worst < 1e-11, worst
    print("OK")
