"""Thin tensor-level wrappers over the C ABI (one function per ``vb_*`` entry point).

Inputs/outputs are CUDA torch tensors; nothing here computes on the host.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _cabi
from ._device import empty, ptr, stream_ptr

COLLAPSE_MODES = {"median": 0, "mean": 1, "sum": 2, "max": 3, "absmean": 4, "wmean": 5, "trimmean": 6}

# upper bound for the derotation scratch planes (bytes); override with VIP_B200_DEROT_SCRATCH
_DEROT_SCRATCH_MAX = int(os.environ.get("VIP_B200_DEROT_SCRATCH", str(8 << 30)))


def _bytes(nbytes, device):
    return torch.empty(int(nbytes), dtype=torch.uint8, device=device)


def gram(M, deflate=False):
    """G (n,n) fp64 = M M^T for a (n,p) fp32 matrix."""
    lib = _cabi.lib()
    n, p = M.shape
    G = empty((n, n), torch.float64, M.device)
    nb = lib.vb_gram_workspace_bytes(n, p)
    ws = _bytes(nb, M.device)
    _cabi.check(lib.vb_gram_f32(ptr(M), n, p, int(bool(deflate)), ptr(G), ptr(ws), nb, stream_ptr()),
                "vb_gram_f32")
    return G


def cross_gram(A, B):
    """C (na,nb) fp64 = A B^T for fp32 matrices sharing the pixel axis."""
    lib = _cabi.lib()
    na, p = A.shape
    nb_, p2 = B.shape
    if p != p2:
        raise ValueError("cross_gram: pixel axes differ")
    Cm = empty((na, nb_), torch.float64, A.device)
    nbytes = lib.vb_cross_gram_workspace_bytes(na, nb_)
    ws = _bytes(nbytes, A.device)
    _cabi.check(lib.vb_cross_gram_f32(ptr(A), na, ptr(B), nb_, p, ptr(Cm), ptr(ws), nbytes, stream_ptr()),
                "vb_cross_gram_f32")
    return Cm


def eigh(G, max_sweeps=0, tol=0.0, check=True):
    """Eigen-decomposition of a symmetric PSD fp64 matrix.

    Returns (evals[n] descending, evecs[n,n] with row j = eigenvector j, info dict).  ``check=False`` skips the
    read-back of the convergence record -- and with it the host synchronisation -- for matrices the single-launch
    solver handles (n <= 128): used for the many small whitening problems of the randomized SVD, whose 30 sweeps
    are far more than a 60 x 60 Gramian ever needs."""
    lib = _cabi.lib()
    n = G.shape[0]
    evals = empty((n,), torch.float64, G.device)
    evecs = empty((n, n), torch.float64, G.device)
    nb = lib.vb_eigh_workspace_bytes(n)
    ws = _bytes(nb, G.device)
    if not check and n <= 128:
        _cabi.check(lib.vb_eigh_f64(ptr(G), n, ptr(evals), ptr(evecs), int(max_sweeps), float(tol), ptr(ws), nb,
                                    None, stream_ptr()), "vb_eigh_f64")
        return evals, evecs, {"sweeps": None, "converged": None}
    info = (C.c_int * 2)()
    _cabi.check(lib.vb_eigh_f64(ptr(G), n, ptr(evals), ptr(evecs), int(max_sweeps), float(tol), ptr(ws), nb,
                                info, stream_ptr()), "vb_eigh_f64")
    if n > 1 and not info[1]:
        raise RuntimeError(f"vip_b200: Jacobi eigensolver did not converge in {info[0]} sweeps")
    return evals, evecs, {"sweeps": int(info[0]), "converged": bool(info[1])}


def chol_whiten(G):
    """Wt (n,n) fp64, lower triangular, with Wt G Wt^T = I for a small SPD Gramian (n <= 128); no host sync."""
    lib = _cabi.lib()
    n = G.shape[0]
    G = G.contiguous()
    Wt = empty((n, n), torch.float64, G.device)
    _cabi.check(lib.vb_chol_whiten_f64(ptr(G), n, ptr(Wt), stream_ptr()), "vb_chol_whiten_f64")
    return Wt


def eigh_topk(G, k, tol=0.0, max_iter=0):
    """Leading k eigenpairs of a symmetric PSD fp64 matrix by block subspace iteration.

    Returns (evals[k] descending, evecs[k,n] rows = eigenvectors, info dict)."""
    lib = _cabi.lib()
    n = G.shape[0]
    evals = empty((k,), torch.float64, G.device)
    evecs = empty((k, n), torch.float64, G.device)
    nb = lib.vb_eigh_topk_workspace_bytes(n, k)
    ws = _bytes(nb, G.device)
    info = (C.c_int * 2)()
    _cabi.check(lib.vb_eigh_topk_f64(ptr(G), n, int(k), float(tol), int(max_iter), ptr(evals), ptr(evecs),
                                     ptr(ws), nb, info, stream_ptr()), "vb_eigh_topk_f64")
    return evals, evecs, {"iters": int(info[0]), "converged": bool(info[1])}


def eigh_topk_async(G, k, tol=0.0, max_iter=0):
    """``eigh_topk`` without the host synchronisation.  Returns (evals, evecs, info) where ``info`` is a pinned
    int32[2] tensor {iterations, converged} that is valid once the current stream has been synchronised.
    The workspace of the cooperative kernel is attached to the record (``info.ws``) so that it lives at least
    as long as the caller holds the record, whatever stream later allocations are made on."""
    lib = _cabi.lib()
    n = G.shape[0]
    evals = empty((k,), torch.float64, G.device)
    evecs = empty((k, n), torch.float64, G.device)
    nb = lib.vb_eigh_topk_workspace_bytes(n, k)
    ws = _bytes(nb, G.device)
    info = torch.zeros(2, dtype=torch.int32).pin_memory()
    _cabi.check(lib.vb_eigh_topk_async_f64(ptr(G), n, int(k), float(tol), int(max_iter), ptr(evals), ptr(evecs),
                                           ptr(ws), nb, info.data_ptr(), stream_ptr()), "vb_eigh_topk_async_f64")
    info.ws = ws
    return evals, evecs, info


def topk_supported(n, k):
    """The subspace solver handles k <= 56 with a block (16, 32 or 64 vectors) no wider than the matrix
    (k <= 24: one fused cooperative launch; 25..56: per-phase kernels, block 64)."""
    return k <= 56 and n >= (16 if k <= 10 else 32 if k <= 24 else 64) and n > 2 * k


def pcs(Wt, M):
    """V (k,p) fp32 = Wt (k,n) . M (n,p); Wt is used in fp64 and the sum is accumulated in fp64."""
    lib = _cabi.lib()
    Wt = Wt.to(torch.float64).contiguous()
    k, n = Wt.shape
    n2, p = M.shape
    assert n == n2
    V = empty((k, p), torch.float32, M.device)
    _cabi.check(lib.vb_pcs_f32(ptr(Wt), ptr(M), k, n, p, ptr(V), stream_ptr()), "vb_pcs_f32")
    return V


def pcs_hilo(Wt, M):
    """(Vhi, Vlo) fp32 (k,p) with Vhi + Vlo = Wt (k,n) . M (n,p) carried to ~48 bits (fp64 accumulation, error-free
    split): for products whose rows span more dynamic range than fp32 holds (raw randomized-SVD sketches)."""
    lib = _cabi.lib()
    Wt = Wt.to(torch.float64).contiguous()
    k, n = Wt.shape
    n2, p = M.shape
    assert n == n2
    Vhi = empty((k, p), torch.float32, M.device)
    Vlo = empty((k, p), torch.float32, M.device)
    _cabi.check(lib.vb_pcs_hilo_f32(ptr(Wt), ptr(M), k, n, p, ptr(Vhi), ptr(Vlo), stream_ptr()), "vb_pcs_hilo_f32")
    return Vhi, Vlo


def project_subtract(M, Cm, V, out=None):
    """R (n,p) = M - Cm (n,k) . V (k,p), fp32."""
    lib = _cabi.lib()
    n, p = M.shape
    k = V.shape[0]
    assert Cm.shape == (n, k) and Cm.is_contiguous()
    R = out if out is not None else empty((n, p), torch.float32, M.device)
    _cabi.check(lib.vb_project_subtract_f32(ptr(M), ptr(Cm), k, ptr(V), k, n, p, ptr(R), stream_ptr()),
                "vb_project_subtract_f32")
    return R


def project_subtract_hp(M, C64, Vhi, Vlo=None, out=None):
    """R (n,p) fp32 = M - C64 (n,k fp64) . (Vhi + Vlo) (k,p), fp64 accumulation, one final rounding."""
    lib = _cabi.lib()
    n, p = M.shape
    k = Vhi.shape[0]
    C64 = C64.to(torch.float64).contiguous()
    assert C64.shape == (n, k) and Vhi.is_contiguous() and (Vlo is None or Vlo.is_contiguous())
    R = out if out is not None else empty((n, p), torch.float32, M.device)
    _cabi.check(lib.vb_project_subtract_hp_f32(ptr(M), ptr(C64), k, ptr(Vhi), ptr(Vlo), k, n, p, ptr(R), stream_ptr()),
                "vb_project_subtract_hp_f32")
    return R


def sub(a, b):
    lib = _cabi.lib()
    out = torch.empty_like(a)
    _cabi.check(lib.vb_sub_f32(ptr(a), ptr(b), ptr(out), a.numel(), stream_ptr()), "vb_sub_f32")
    return out


def derotate(cube, krot, a, b, S, N, y0, mask_val=float("nan"), zero_masked=False, force_direct=False,
             scratch_max=None):
    """Rotate every frame of a (n,S,S) fp32 CUDA cube (vip-fft semantics); per-frame host scalars
    krot/a/b come from ``preproc.derotation.rotation_scalars``."""
    lib = _cabi.lib()
    n = cube.shape[0]
    dev = cube.device
    out = torch.empty_like(cube)
    d_k = torch.as_tensor(np.asarray(krot, dtype=np.int32)).to(dev)
    d_a = torch.as_tensor(np.asarray(a, dtype=np.float64)).to(dev)
    d_b = torch.as_tensor(np.asarray(b, dtype=np.float64)).to(dev)
    nbytes = lib.vb_derotate_scratch_bytes(n, S, N, scratch_max or _DEROT_SCRATCH_MAX)
    scratch = _bytes(nbytes, dev)
    mask_is_nan = int(np.isnan(mask_val))
    _cabi.check(lib.vb_derotate_f32(ptr(cube), ptr(out), n, S, N, y0, ptr(d_k), ptr(d_a), ptr(d_b),
                                    float(mask_val), mask_is_nan, int(bool(zero_masked)), ptr(scratch),
                                    nbytes, int(bool(force_direct)), stream_ptr()), "vb_derotate_f32")
    return out


def derotate_scatter(cube, krot, a, b, S, N, y0, out_bases, rows_per_shard, frame_stride, frame_offset,
                     mask_val=float("nan"), zero_masked=False, scratch_max=None):
    """``derotate`` whose output rows go straight into the pixel-shard slabs ``out_bases`` (list of device pointers,
    local or peer; see ``vb_derotate_scatter_f32``).  Returns nothing: the slabs are the output."""
    lib = _cabi.lib()
    n = cube.shape[0]
    dev = cube.device
    d_k = torch.as_tensor(np.asarray(krot, dtype=np.int32)).to(dev)
    d_a = torch.as_tensor(np.asarray(a, dtype=np.float64)).to(dev)
    d_b = torch.as_tensor(np.asarray(b, dtype=np.float64)).to(dev)
    nbytes = lib.vb_derotate_scratch_bytes(n, S, N, scratch_max or _DEROT_SCRATCH_MAX)
    scratch = _bytes(nbytes, dev)
    bases = (C.c_void_p * len(out_bases))(*[int(x) for x in out_bases])
    _cabi.check(lib.vb_derotate_scatter_f32(ptr(cube), n, S, N, y0, ptr(d_k), ptr(d_a), ptr(d_b), float(mask_val),
                                            int(np.isnan(mask_val)), int(bool(zero_masked)), ptr(scratch), nbytes,
                                            bases, len(out_bases), int(rows_per_shard), int(frame_stride),
                                            int(frame_offset), stream_ptr()), "vb_derotate_scatter_f32")


def project_subtract_hp_rows(M, C64, Vhi, Vlo, row_ptrs, scratch=None):
    """``project_subtract_hp`` with row i of the result written to the address ``row_ptrs[i]`` (int64 CUDA tensor of
    n device pointers, local or peer memory).  ``scratch`` (n,p) fp32 is needed when k > 32."""
    lib = _cabi.lib()
    n, p = M.shape
    k = Vhi.shape[0]
    C64 = C64.to(torch.float64).contiguous()
    assert C64.shape == (n, k) and row_ptrs.dtype == torch.int64 and row_ptrs.numel() == n
    if k > 32 and scratch is None:
        scratch = empty((n, p), torch.float32, M.device)
    _cabi.check(lib.vb_project_subtract_hp_rows_f32(ptr(M), ptr(C64), k, ptr(Vhi), ptr(Vlo), k, n, p, ptr(scratch),
                                                    ptr(row_ptrs), stream_ptr()), "vb_project_subtract_hp_rows_f32")


def collapse(cube2d, mode="median", w=None, trim_k=0, trim_n=0):
    """(n,p) fp32 -> (p,) fp32 (fp64 for 'wmean')."""
    lib = _cabi.lib()
    n, p = cube2d.shape
    m = COLLAPSE_MODES[mode]
    dev = cube2d.device
    d_w = None
    if mode == "wmean":
        d_w = torch.as_tensor(np.asarray(w, dtype=np.float64)).to(dev)
        out = empty((p,), torch.float64, dev)
    else:
        out = empty((p,), torch.float32, dev)
    _cabi.check(lib.vb_collapse_f32(ptr(cube2d), n, p, m, ptr(d_w), int(trim_k), int(trim_n), ptr(out),
                                    stream_ptr()), "vb_collapse_f32")
    return out


def gather_columns(M, cols):
    """(n,p) fp32, int32 column indices (npx,) -> (n,npx) fp32."""
    lib = _cabi.lib()
    n, p = M.shape
    npx = cols.numel()
    out = empty((n, npx), torch.float32, M.device)
    _cabi.check(lib.vb_gather_columns_f32(ptr(M), n, p, ptr(cols), npx, ptr(out), stream_ptr()),
                "vb_gather_columns_f32")
    return out


def scatter_columns(src, cols, dst):
    """dst[:, cols] = src  for (n,npx) src and (n,p) dst, in place."""
    lib = _cabi.lib()
    n, npx = src.shape
    _cabi.check(lib.vb_scatter_columns_f32(ptr(src), n, npx, ptr(cols), dst.shape[1], ptr(dst), stream_ptr()),
                "vb_scatter_columns_f32")
    return dst


def annular_weights(G, idx, lens, frames, ncomp, tol=0.0, max_iter=40, direct_fallback=True, force_direct=False):
    """Per-problem projection weights from the library Gramian G (nlib,nlib) fp64.

    idx (nprob,Lmax) int32, lens (nprob,) int32, frames (nprob,) int32 = row of G of each target.
    Block subspace iteration first (``max_iter`` steps); problems it leaves unconverged (flat,
    noise-dominated spectra) are solved by the direct tridiagonal solver.
    Returns (W (nprob,nlib) fp32, iters (nprob,) int32: >0 iterations, 100000 = direct, <0 = failed)."""
    lib = _cabi.lib()
    n = G.shape[0]
    nprob, Lmax = idx.shape
    dev = G.device
    W = torch.zeros((nprob, n), dtype=torch.float32, device=dev)
    iters = torch.zeros((nprob,), dtype=torch.int32, device=dev)
    if not force_direct:
        _cabi.check(lib.vb_annular_weights_f64(ptr(G), 0, n, ptr(idx), ptr(lens), ptr(frames), nprob, Lmax,
                                               int(ncomp), float(tol), int(max_iter), ptr(W), ptr(iters),
                                               stream_ptr()), "vb_annular_weights_f64")
        if not direct_fallback:
            return W, iters
        todo = torch.nonzero(iters < 0).flatten().to(torch.int32)
    else:
        todo = torch.arange(nprob, dtype=torch.int32, device=dev)
    chunk = max(1, min(1024, (1 << 30) // (Lmax * Lmax * 8)))      # <= 1 GiB of workspace per launch
    for s in range(0, todo.numel(), chunk):
        part = todo[s:s + chunk].contiguous()
        ws = torch.empty(part.numel() * Lmax * Lmax, dtype=torch.float64, device=dev)
        _cabi.check(lib.vb_annular_direct_f64(ptr(G), 0, n, ptr(idx), ptr(lens), ptr(frames), nprob, Lmax,
                                              int(ncomp), ptr(part), part.numel(), ptr(W), ptr(iters), ptr(ws),
                                              stream_ptr()), "vb_annular_direct_f64")
    return W, iters


def annular_weights_auto(G, idx, lens, frames, rowsum, npx, noise_tol, kmax=24):
    """``annular_weights`` with the number of components of every problem chosen by the reference's noise-decay rule
    (``get_eigenvectors(ncomp='auto')``, ``psfsub/svd.py:622-672``) among ``kmax`` eigenpairs (direct solver).
    ``rowsum`` (nlib,) fp64 = sum over the pixels of every row of the library matrix, ``npx`` its width.
    Returns (W (nprob,nlib) fp32, ncomp (nprob,) int32; negative = clipped at kmax)."""
    lib = _cabi.lib()
    n = G.shape[0]
    nprob, Lmax = idx.shape
    dev = G.device
    W = torch.zeros((nprob, n), dtype=torch.float32, device=dev)
    iters = torch.zeros((nprob,), dtype=torch.int32, device=dev)
    ncomp = torch.zeros((nprob,), dtype=torch.int32, device=dev)
    rowsum = rowsum.to(torch.float64).contiguous()
    todo = torch.arange(nprob, dtype=torch.int32, device=dev)
    chunk = max(1, min(1024, (1 << 30) // (Lmax * Lmax * 8)))
    for s in range(0, nprob, chunk):
        part = todo[s:s + chunk].contiguous()
        ws = torch.empty(part.numel() * Lmax * Lmax, dtype=torch.float64, device=dev)
        _cabi.check(lib.vb_annular_auto_f64(ptr(G), 0, n, ptr(idx), ptr(lens), ptr(frames), nprob, Lmax, int(kmax),
                                            ptr(rowsum), float(npx), float(noise_tol), ptr(part), part.numel(),
                                            ptr(W), ptr(iters), ptr(ncomp), ptr(ws), stream_ptr()),
                    "vb_annular_auto_f64")
    return W, ncomp


def upload_columns(host2d, c0, c1, device):
    """Columns [c0, c1) of a C-contiguous fp32 host matrix -> contiguous (n, c1-c0) CUDA tensor (one 2-D DMA)."""
    lib = _cabi.lib()
    assert host2d.dtype == np.float32 and host2d.flags["C_CONTIGUOUS"] and host2d.ndim == 2
    n, p = host2d.shape
    out = empty((n, c1 - c0), torch.float32, device)
    src = host2d.ctypes.data + c0 * 4
    _cabi.check(lib.vb_memcpy2d_h2d(ptr(out), (c1 - c0) * 4, src, p * 4, (c1 - c0) * 4, n, stream_ptr()),
                "vb_memcpy2d_h2d")
    return out


def gemm(A, B, C, trans_b=False, alpha=1.0, beta=0.0, a_mod=0, b_mod=0):
    """Batched fp32 GEMM on 3-d tensors: C[b] = alpha * A[b % a_mod or b] @ op(B[b % b_mod or b]) + beta * C[b].

    A (Ba,M,K), B (Bb,K,N) or (Bb,N,K) if trans_b, C (batch,M,N); the last dimension of each must be
    contiguous (row strides are passed as leading dimensions, so column-block views are fine)."""
    lib = _cabi.lib()
    batch, M, N = C.shape
    K = A.shape[2]
    for t in (A, B, C):
        assert t.dtype == torch.float32 and t.stride(2) == 1
    _cabi.check(lib.vb_gemm_f32(ptr(A), A.stride(1), A.stride(0), int(a_mod), ptr(B), B.stride(1), B.stride(0),
                                int(b_mod), int(bool(trans_b)), ptr(C), C.stride(1), C.stride(0), M, N, K,
                                float(alpha), float(beta), batch, stream_ptr()), "vb_gemm_f32")
    return C


class Planes3:
    """fp32 matrix (rows, K) as three bf16 planes (3, rows, ldp) -- the K-major operand format of the tcgen05 kernels."""

    def __init__(self, data, rows, K):
        self.data, self.rows, self.K = data, int(rows), int(K)

    @property
    def ldp(self):
        return int(self.data.shape[2])

    @property
    def plane_stride(self):
        return int(self.data.stride(0))


def planes3_empty(rows, K, device):
    ldp = (int(K) + 7) // 8 * 8
    return Planes3(torch.empty((3, int(rows), ldp), dtype=torch.bfloat16, device=device), rows, K)


def split3(X):
    """Error-free split of a contiguous fp32 matrix (rows, K) into three bf16 planes (``vb_split3_bf16``)."""
    lib = _cabi.lib()
    assert X.dtype == torch.float32 and X.dim() == 2 and X.stride(1) == 1
    rows, K = X.shape
    P = planes3_empty(rows, K, X.device)
    _cabi.check(lib.vb_split3_bf16(ptr(X), rows, K, X.stride(0), ptr(P.data), P.ldp, P.plane_stride, stream_ptr()),
                "vb_split3_bf16")
    return P


def gemm_tc(A, a_mod, B, M, N, batch, out=None, out_planes=None, msplit=0):
    """Batched ``C[b] = A[b % a_mod] @ B[b]^T`` on the tensor cores (``vb_gemm_bf16x3_tc``): A, B are :class:`Planes3`
    with a_mod * M and batch * N rows and the same K.  ``out`` (batch, M, N) fp32, or ``out_planes`` (:class:`Planes3`
    with batch * msplit rows): the bf16x3 planes of C, rows r >= msplit folded to row r - msplit at column offset N."""
    lib = _cabi.lib()
    assert A.K == B.K and A.rows == a_mod * M and B.rows == batch * N
    if out_planes is not None:
        _cabi.check(lib.vb_gemm_bf16x3_tc(ptr(A.data), A.ldp, A.plane_stride, int(a_mod), ptr(B.data), B.ldp,
                                          B.plane_stride, M, N, A.K, batch, 0, 0, 0, ptr(out_planes.data),
                                          out_planes.ldp, out_planes.plane_stride, int(msplit), stream_ptr()),
                    "vb_gemm_bf16x3_tc")
        return out_planes
    assert out.dtype == torch.float32 and out.is_contiguous() and tuple(out.shape) == (batch, M, N)
    _cabi.check(lib.vb_gemm_bf16x3_tc(ptr(A.data), A.ldp, A.plane_stride, int(a_mod), ptr(B.data), B.ldp,
                                      B.plane_stride, M, N, A.K, batch, ptr(out), N, M * N, 0, 0, 0, 0,
                                      stream_ptr()), "vb_gemm_bf16x3_tc")
    return out


def shift_operators(shifts, nplane, L, device):
    """(nf, L, L) fp32 Toeplitz operators of the vip-fft shift: T[f][m][n] = Re D_N(m - n - shift[f])."""
    lib = _cabi.lib()
    nf = len(shifts)
    d_s = torch.as_tensor(np.asarray(shifts, dtype=np.float64)).to(device)
    d_n = torch.as_tensor(np.asarray(nplane, dtype=np.int32)).to(device)
    T = empty((nf, L, L), torch.float32, device)
    _cabi.check(lib.vb_shift_operators_f32(ptr(d_s), ptr(d_n), nf, L, ptr(T), stream_ptr()), "vb_shift_operators_f32")
    return T


def checker_correct(X, out, coef):
    """out[f] -= (-1)^(r+c) coef[f] sum_{r',c'} (-1)^(r'+c') X[f]  (Nyquist term of the separable shift)."""
    lib = _cabi.lib()
    nf, ny, nx = X.shape
    assert X.is_contiguous() and out.is_contiguous()
    d_c = torch.as_tensor(np.asarray(coef, dtype=np.float64)).to(X.device)
    ws = empty((nf,), torch.float64, X.device)
    _cabi.check(lib.vb_checker_correct_f32(ptr(X), ptr(out), nf, ny, nx, ptr(d_c), ptr(ws), stream_ptr()),
                "vb_checker_correct_f32")
    return out


def upload_and_gram(host2d, device, nslabs=8):
    """C-contiguous fp32 host matrix (n,p) -> (M on the device, G = M M^T fp64), upload and SYRK pipelined."""
    lib = _cabi.lib()
    assert host2d.dtype == np.float32 and host2d.flags["C_CONTIGUOUS"] and host2d.ndim == 2
    n, p = host2d.shape
    M = empty((n, p), torch.float32, device)
    G = empty((n, n), torch.float64, device)
    nb = lib.vb_gram_workspace_bytes(n, p)
    ws = _bytes(nb, device)
    _cabi.check(lib.vb_upload_gram_f32(host2d.ctypes.data, n, p, ptr(M), ptr(G), ptr(ws), nb, int(nslabs),
                                       stream_ptr()), "vb_upload_gram_f32")
    return M, G
