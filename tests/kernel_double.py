"""TEST INFRASTRUCTURE: a CPU stand-in for the ``vb_*`` kernels, so that the HOST orchestration of the drop-in
callables (argument parsing, branch selection, library bookkeeping, return layouts) can be exercised in the build
container, which has no GPU.  Every function mirrors the contract of its namesake in ``vip_b200/kernels.py`` with
plain torch-CPU / numpy arithmetic (fp64 where the kernels accumulate in fp64) or the oracle's restatement.

Only tests install it (``install(monkeypatch)``); the product never imports this module and still refuses to run
without CUDA (``tests/test_host_logic.py::test_no_cpu_fallback_without_cuda``).  GPU parity of the real kernels is
``tests/test_gpu_parity.py``.
"""
import numpy as np
import torch

from oracle import vip_oracle as O

CPU = torch.device("cpu")


def gram(M, deflate=False):
    Md = M.double()
    return Md @ Md.T


def cross_gram(A, B):
    return A.double() @ B.double().T


def eigh(G, max_sweeps=0, tol=0.0, check=True):
    w, v = torch.linalg.eigh(G.double())
    w = w.abs()                                   # the one-sided Jacobi kernel returns column norms
    order = torch.argsort(w, descending=True, stable=True)
    return w[order].contiguous(), v[:, order].T.contiguous(), {"sweeps": 1, "converged": True}


def chol_whiten(G):
    g = G.double().numpy()
    g = 0.5 * (g + g.T)
    R = np.linalg.cholesky(g).T                       # G = R^T R
    return torch.from_numpy(np.ascontiguousarray(np.linalg.inv(R).T))


def eigh_topk(G, k, tol=0.0, max_iter=0):
    w, E, _ = eigh(G)
    return w[:k].contiguous(), E[:k].contiguous(), {"iters": 1, "converged": True}


def topk_supported(n, k):
    return False                                   # always the synchronous full solver: no CUDA stream needed


def pcs(Wt, M):
    return (Wt.double() @ M.double()).float()


def pcs_hilo(Wt, M):
    acc = Wt.numpy().astype(np.float64) @ M.numpy().astype(np.float64)
    hi = acc.astype(np.float32)
    return torch.from_numpy(hi), torch.from_numpy((acc - hi).astype(np.float32))


def project_subtract(M, Cm, V, out=None):
    R = M - Cm.float() @ V
    if out is not None:
        out.copy_(R)
        return out
    return R


def project_subtract_hp(M, C64, Vhi, Vlo=None, out=None):
    V = Vhi.double() if Vlo is None else Vhi.double() + Vlo.double()
    R = (M.double() - C64.double() @ V).float()
    if out is not None:
        out.copy_(R)
        return out
    return R


def sub(a, b):
    return a - b


def derotate(cube, krot, a, b, S, N, y0, mask_val=float("nan"), zero_masked=False, force_direct=False,
             scratch_max=None):
    """Rotation from the per-frame scalars: angle = 90 krot + residual, residual = -asin(b) (b = -sin(residual))."""
    arr = cube.numpy()
    out = np.empty_like(arr)
    for i in range(arr.shape[0]):
        angle = 90.0 * int(krot[i]) + float(np.rad2deg(-np.arcsin(b[i])))
        out[i] = O.frame_rotate(arr[i], angle, mask_val=mask_val, interp_zeros=bool(zero_masked))
    return torch.from_numpy(out)


def collapse(cube2d, mode="median", w=None, trim_k=0, trim_n=0):
    import warnings
    a = cube2d.numpy()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if mode == "median":
            r = np.nanmedian(a, axis=0)
        elif mode == "mean":
            r = np.nanmean(a, axis=0)
        elif mode == "sum":
            r = np.nansum(a, axis=0)
        elif mode == "max":
            r = np.nanmax(a, axis=0)
        elif mode == "absmean":
            r = np.nanmean(np.abs(a), axis=0)
        elif mode == "wmean":
            r = np.inner(np.asarray(w, dtype=np.float64), np.moveaxis(np.nan_to_num(a).astype(np.float64), 0, -1))
        elif mode == "trimmean":
            r = np.nanmean(np.sort(a, axis=0)[trim_k:trim_k + trim_n], axis=0)
        else:
            raise KeyError(mode)
    return torch.from_numpy(np.ascontiguousarray(r))


def upload_and_gram(host2d, device, nslabs=8):
    M = torch.from_numpy(np.array(host2d, dtype=np.float32, copy=True))
    return M, gram(M)


def upload_columns(host2d, c0, c1, device):
    return torch.from_numpy(np.ascontiguousarray(host2d[:, c0:c1]))


def gather_columns(M, cols):
    return M[:, cols.long()].contiguous()


def scatter_columns(src, cols, dst):
    dst[:, cols.long()] = src
    return dst


def gemm(A, B, C, trans_b=False, alpha=1.0, beta=0.0, a_mod=0, b_mod=0):
    for i in range(C.shape[0]):
        a = A[i % a_mod if a_mod else i]
        b = B[i % b_mod if b_mod else i]
        prod = a.double() @ (b.double().T if trans_b else b.double())
        C[i] = (alpha * prod + (beta * C[i].double() if beta != 0.0 else 0.0)).float()   # beta = 0: C is not read
    return C


def split3(X):
    """Contract of ``kernels.split3``: exact three-way bf16 split, planes (3, rows, ldp) with ldp = K rounded up to 8."""
    from vip_b200.kernels import Planes3
    rows, K = X.shape
    ldp = (K + 7) // 8 * 8
    data = torch.zeros((3, rows, ldp), dtype=torch.bfloat16)
    r = X.float().clone()
    for q in range(3):
        h = r.to(torch.bfloat16)
        data[q, :, :K] = h
        r = r - h.float()
    return Planes3(data, rows, K)


def planes3_empty(rows, K, device):
    from vip_b200.kernels import Planes3
    return Planes3(torch.zeros((3, int(rows), (int(K) + 7) // 8 * 8), dtype=torch.bfloat16), rows, K)


def gemm_tc(A, a_mod, B, M, N, batch, out=None, out_planes=None, msplit=0):
    """Contract of ``kernels.gemm_tc`` (products in float64 of the summed planes)."""
    Af = A.data.double().sum(0)[:, :A.K].reshape(a_mod, M, A.K)
    Bf = B.data.double().sum(0)[:, :B.K].reshape(batch, N, B.K)
    for b in range(batch):
        Cb = (Af[b % a_mod] @ Bf[b].T).float()                  # (M, N)
        if out_planes is None:
            out[b] = Cb
            continue
        folded = torch.cat((Cb[:msplit], Cb[msplit:]), dim=1) if M > msplit else Cb      # (msplit, 2N)
        r = folded.clone()
        for q in range(3):
            h = r.to(torch.bfloat16)
            out_planes.data[q, b * msplit:(b + 1) * msplit, :folded.shape[1]] = h
            r = r - h.float()
    return out if out_planes is None else out_planes


def annular_weights(G, idx, lens, frames, ncomp, tol=0.0, max_iter=40, direct_fallback=True, force_direct=False):
    """Per-problem projection weights from the library Gramian (contract of ``kernels.annular_weights``):
    W[q, I] = X diag(1/theta) X^T G[I, frame_q] with (theta, X) the leading eigenpairs of G[I, I]."""
    Gn = G.double().numpy()
    nprob = idx.shape[0]
    W = np.zeros((nprob, Gn.shape[0]), dtype=np.float32)
    for q in range(nprob):
        I = idx[q, :int(lens[q])].numpy().astype(np.int64)
        w_, v_ = np.linalg.eigh(Gn[np.ix_(I, I)])
        kk = min(int(ncomp), len(I))
        X, th = v_[:, ::-1][:, :kk], w_[::-1][:kk]
        W[q, I] = X @ ((X.T @ Gn[I, int(frames[q])]) / th)
    return torch.from_numpy(W), torch.ones(nprob, dtype=torch.int32)


def annular_weights_auto(G, idx, lens, frames, rowsum, npx, noise_tol, kmax=24):
    """Contract of ``kernels.annular_weights_auto``: the noise-decay rule evaluated from the eigenpairs of the
    library Gramian and the row sums of the library matrix (no access to the matrix itself, like the kernel)."""
    Gn = G.double().numpy()
    rs = rowsum.double().numpy()
    nprob = idx.shape[0]
    W = np.zeros((nprob, Gn.shape[0]), dtype=np.float32)
    used = np.zeros(nprob, dtype=np.int32)
    for q in range(nprob):
        I = idx[q, :int(lens[q])].numpy().astype(np.int64)
        L = len(I)
        w_, v_ = np.linalg.eigh(Gn[np.ix_(I, I)])
        lam, X = w_[::-1], v_[:, ::-1]
        tot = L * float(npx)
        max_evs = int(min(L, npx))
        k = min(int(kmax), L)
        s2, sr, prev, decay, m, clipped = np.trace(Gn[np.ix_(I, I)]), rs[I].sum(), 0.0, 1.0, 0, False
        while decay >= noise_tol:
            m += 1
            if m <= max_evs:
                if m > k:
                    clipped = True
                    break
                s2 -= lam[m - 1]
                sr -= X[:, m - 1].sum() * (X[:, m - 1] @ rs[I])
            noise = np.sqrt(max(s2 / tot - (sr / tot) ** 2, 0.0))
            if m > 1:
                decay = prev - noise
            prev = noise
        m = k if clipped else min(m, max_evs)
        used[q] = -m if clipped else m
        Xm, th = X[:, :m], lam[:m]
        W[q, I] = Xm @ ((Xm.T @ Gn[I, int(frames[q])]) / th)
    return torch.from_numpy(W), torch.from_numpy(used)


_NAMES = ("annular_weights_auto", "annular_weights", "gram", "cross_gram", "eigh", "chol_whiten", "eigh_topk", "topk_supported", "pcs", "pcs_hilo", "project_subtract", "project_subtract_hp", "sub", "derotate",
          "collapse", "upload_and_gram", "upload_columns", "gather_columns", "scatter_columns", "gemm", "split3", "planes3_empty", "gemm_tc")


def aperture_sums_device(frame_dev, xs, ys, r):
    from oracle import vip_oracle as O
    return torch.from_numpy(O.aperture_sums_exact(frame_dev.numpy().astype(np.float64), xs, ys, r))


def snr_points_device(frame_dev, xs, ys, fwhm, frame2_dev=None, use2alone=False, exclude_negative_lobes=False):
    from oracle import vip_oracle as O
    a = frame_dev.numpy()
    b = None if frame2_dev is None else frame2_dev.numpy()
    res = [O.snr(a, (int(x), int(y)), fwhm, True, b, use2alone, exclude_negative_lobes) for x, y in zip(xs, ys)]
    return (torch.tensor([r[-1] for r in res], dtype=torch.float64),
            torch.tensor([r[2] for r in res], dtype=torch.float64))


def local_max_mask_device(frame_dev, min_distance, threshold):
    from oracle import vip_oracle as O
    return torch.from_numpy(O.local_max_mask(frame_dev.numpy(), int(min_distance), float(threshold)))


def install(monkeypatch):
    """Route ``vip_b200`` through the stand-ins above for the duration of one test."""
    import vip_b200
    from vip_b200 import kernels, _device
    from vip_b200.psfsub import pca_fullfr, annular, sdi
    g = globals()
    for name in _NAMES:
        monkeypatch.setattr(kernels, name, g[name])
    monkeypatch.setattr(_device, "require_cuda", lambda: CPU)
    monkeypatch.setattr(_device, "free_memory_bytes", lambda: 8 << 30)
    from vip_b200.metrics import snr_source
    monkeypatch.setattr(snr_source, "aperture_sums_device", aperture_sums_device)
    monkeypatch.setattr(snr_source, "snr_points_device", snr_points_device)
    import importlib
    detection_mod = importlib.import_module("vip_b200.metrics.detection")
    monkeypatch.setattr(detection_mod, "local_max_mask_device", local_max_mask_device)
    monkeypatch.setattr(detection_mod, "require_cuda", lambda: CPU)
    for mod in (pca_fullfr, annular, sdi, snr_source):
        if hasattr(mod, "require_cuda"):
            monkeypatch.setattr(mod, "require_cuda", lambda: CPU)
    return vip_b200
