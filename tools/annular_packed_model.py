"""Executable numpy model of the PACKED direct solver of csrc/annular.cu (annular_direct_one<true>): same packed
indexing, same loop structure per 'thread', run serially.  Checks (1) the tridiagonal matrix has the spectrum of the
input, (2) the LU inverse iteration + back-transformation return eigenvectors of the input.  Build-container check
of the index arithmetic (there is no GPU here); the GPU parity tests are tests/test_gpu_parity.py -k annular."""
import numpy as np


def off(c, L):
    return c * L - c * (c - 1) // 2


def pack(G):
    L = G.shape[0]
    A = np.zeros(L * (L + 1) // 2)
    for c in range(L):
        col = off(c, L) - c
        for r in range(c, L):
            A[col + r] = G[r, c]
    return A


def tridiag(A, L):
    d = np.zeros(L); e = np.zeros(L); tau = np.zeros(L)
    for j in range(L - 1):
        m = L - 1 - j
        cj = off(j, L)
        x = cj + 1
        sigma = sum(A[x + i] ** 2 for i in range(1, m))
        alpha = A[x]
        d[j] = A[cj]
        if sigma == 0.0:
            tau[j] = 0.0; e[j] = alpha
            A[x] = 1.0
            continue
        beta = -np.copysign(np.sqrt(alpha * alpha + sigma), alpha)
        tau[j] = (beta - alpha) / beta
        e[j] = beta
        tj = tau[j]; scale = 1.0 / (alpha - beta)
        vv = np.zeros(m)
        for i in range(m):
            v = 1.0 if i == 0 else A[x + i] * scale
            vv[i] = v; A[x + i] = v
        B = cj + (m + 1)
        pp = np.zeros(m)
        for i in range(m):                      # 'thread' i
            own = B + i * m - i * (i - 1) // 2 - i
            acc = 0.0
            head = i
            for l in range(m):
                b = A[B + head] if l <= i else A[own + l]
                head += m - 1 - l
                acc += b * vv[l]
            pp[i] = tj * acc
        K = -0.5 * tj * float(pp @ vv)
        pp = pp + K * vv
        nrb = (m + 31) >> 5
        for rb in range(nrb):
            for lane in range(32):
                r = (rb << 5) + lane
                if r >= m:
                    continue
                for cbk in range(rb + 1):
                    c0 = cbk << 5
                    c1 = min(c0 + 31, r)
                    el = B + c0 * m - c0 * (c0 - 1) // 2 + (r - c0)
                    for c in range(c0, c1 + 1):
                        A[el] -= vv[c] * pp[r] + pp[c] * vv[r]
                        el += m - 1 - c
    d[L - 1] = A[off(L - 1, L)]
    return d, e, tau


def inverse_iteration(d, e, lam, L, k, rounds=3, seed=0):
    rng = np.random.default_rng(seed)
    Z = rng.uniform(-1, 1, (k, L))
    for _ in range(rounds):
        for r in range(k):
            b = Z[r]
            UF = np.zeros((L, 3))
            lr = lam[r]
            diag = d[0] - lr
            sup = e[0] if L > 1 else 0.0
            bi = b[0]
            for i in range(L - 1):
                bn = b[i + 1]
                sub = e[i]; nd = d[i + 1] - lr; ns = e[i + 1] if i + 2 < L else 0.0
                if abs(diag) >= abs(sub) or abs(sub) < 1e-300:
                    piv = diag if abs(diag) >= 1e-300 else np.copysign(1e-300, diag)
                    rinv = 1.0 / piv
                    mlt = sub * rinv
                    UF[i] = (rinv, sup, 0.0)
                    b[i] = bi
                    bi = bn - mlt * bi
                    diag = nd - mlt * sup
                    sup = ns
                else:
                    rinv = 1.0 / sub
                    mlt = diag * rinv
                    UF[i] = (rinv, nd, ns)
                    b[i] = bn
                    bi = bi - mlt * bn
                    diag = sup - mlt * nd
                    sup = -mlt * ns
            piv = diag if abs(diag) >= 1e-300 else np.copysign(1e-300, diag)
            UF[L - 1] = (1.0 / piv, 0.0, 0.0)
            b[L - 1] = bi
            x1 = x2 = 0.0; mx = 0.0
            PF = 4
            for i0 in range(L - 1, -1, -PF):
                for t in range(PF):
                    i = i0 - t
                    if i >= 0:
                        xi = (b[i] - UF[i, 1] * x1 - UF[i, 2] * x2) * UF[i, 0]
                        b[i] = xi; mx = max(mx, abs(xi)); x2 = x1; x1 = xi
            b *= 1.0 / mx if mx > 0 else 1.0
        for r in range(k):                       # modified Gram-Schmidt
            for s in range(r):
                Z[r] -= (Z[r] @ Z[s]) * Z[s]
            Z[r] /= np.linalg.norm(Z[r])
    return Z


def back_transform(A, tau, Z, L):
    for z in Z:
        for j in range(L - 2, -1, -1):
            tj = tau[j]
            if tj == 0.0:
                continue
            m = L - 1 - j
            v = A[off(j, L) + 1: off(j, L) + 1 + m]
            dt = float(v @ z[j + 1: j + 1 + m]) * tj
            z[j + 1: j + 1 + m] -= dt * v
    return Z


def check(L, k, seed):
    rng = np.random.default_rng(seed)
    M = rng.normal(size=(L, 3 * L)) + 5.0 * rng.normal(size=(L, 1))
    G = M @ M.T
    A = pack(G)
    d, e, tau = tridiag(A, L)
    T = np.diag(d) + np.diag(e[:L - 1], 1) + np.diag(e[:L - 1], -1)
    wT = np.linalg.eigvalsh(T); wG = np.linalg.eigvalsh(G)
    err_spec = np.max(np.abs(wT - wG)) / wG[-1]
    lam = wT[::-1][:k]
    Z = inverse_iteration(d, e, lam, L, k)
    X = back_transform(A, tau, Z, L)
    res = max(np.linalg.norm(G @ X[r] - lam[r] * X[r]) / lam[0] for r in range(k))
    orth = np.max(np.abs(X @ X.T - np.eye(k)))
    return err_spec, res, orth


if __name__ == "__main__":
    for (L, k, seed) in [(1, 1, 0), (2, 2, 1), (5, 3, 2), (33, 10, 3), (64, 10, 4), (97, 24, 5), (130, 10, 6)]:
        es, res, orth = check(L, k, seed)
        print(f"L={L:4d} k={k:3d}  spectrum {es:.2e}  residual {res:.2e}  orthogonality {orth:.2e}")
        assert es < 1e-13 and res < 1e-12 and orth < 1e-12
    print("packed direct-solver model: ok")
