"""Truncated SVD of the frames x pixels matrix on the GPU (``vip_hci/psfsub/svd.py:342-620``).

All deterministic reference back-ends ('lapack', 'eigen', 'arpack' and their cupy/pytorch
variants) return the same top-``ncomp`` right singular vectors up to sign; they are served by one
path: fp64 Gramian of the matrix -> fp64 Jacobi eigensolver -> PCs by a skinny
GEMM.  'randsvd' follows scikit-learn's ``randomized_svd`` (see ``randomized_pcs``).
"""
import numpy as np
import torch

from .. import kernels

_EXACT_MODES = ("lapack", "eigen", "arpack", "cupy", "eigencupy", "pytorch", "eigenpytorch")
_RAND_MODES = ("randsvd", "randcupy", "randpytorch")


def _mode_name(mode):
    return str(getattr(mode, "value", mode))


class Decomposition:
    """Eigen-decomposition of M M^T for a device matrix M (n,p): singular values and left vectors.

    ``ncomp=None`` -> full spectrum (Jacobi; needed for CEVR); an integer ``ncomp`` small enough for
    the subspace solver -> only the leading pairs (falls back to Jacobi if it does not converge)."""

    def __init__(self, M, ncomp=None, G=None, pending=None):
        """``pending``: a list.  When given and the subspace solver applies, it runs WITHOUT the host
        synchronisation of its convergence check: the pinned {iterations, converged} record is appended to
        ``pending`` and the caller, after enqueueing the rest of its pipeline and synchronising once at the
        end, must verify ``converged`` (and redo the work with ``pending=None`` if it is 0)."""
        self.M = M
        n = M.shape[0]
        if G is None:
            G = kernels.gram(M)
        self.full = True
        if ncomp is not None and kernels.topk_supported(n, ncomp):
            if pending is not None:
                evals, evecs, rec = kernels.eigh_topk_async(G, ncomp)
                pending.append(rec)
                self.info = {"deferred": True}
                self.full = False
            else:
                evals, evecs, self.info = kernels.eigh_topk(G, ncomp)
                self.full = not self.info["converged"]
        if self.full:
            evals, evecs, self.info = kernels.eigh(G)
        self.evals = evals                     # descending, fp64 (all n, or the leading ncomp)
        self.U = evecs                         # fp64, row j = j-th left singular vector of M
        self.S = torch.sqrt(torch.clamp(evals, min=0.0))

    def check_rank(self, ncomp):
        n, p = self.M.shape
        if ncomp > min(n, p):
            msg = "{} PCs cannot be obtained from a matrix with size [{},{}]."
            msg += " Increase the size of the patches or request less PCs"
            raise RuntimeError(msg.format(ncomp, n, p))

    def pcs(self, ncomp):
        """V (ncomp,p) fp32: right singular vectors = diag(1/s) U_k M."""
        self.check_rank(ncomp)
        s = self.S[:ncomp]
        Wt = (self.U[:ncomp] / s[:, None]).contiguous()          # fp64: see csrc/proj.cu
        return kernels.pcs(Wt, self.M)

    def pcs_hilo(self, ncomp):
        """(Vhi, Vlo): the same components carried as an error-free fp32 pair (~48 bits) for the high-precision
        projection (``kernels.project_subtract_hp``); Vhi is what ``pcs`` returns."""
        self.check_rank(ncomp)
        s = self.S[:ncomp]
        Wt = (self.U[:ncomp] / s[:, None]).contiguous()
        return kernels.pcs_hilo(Wt, self.M)

    def coeffs64(self, ncomp):
        """C (n,ncomp) fp64 = U_k^T diag(s) (see ``coeffs``)."""
        return (self.U[:ncomp] * self.S[:ncomp, None]).t().contiguous()

    def coeffs(self, ncomp):
        """C (n,ncomp) fp32 = M V^T = U_k^T diag(s): projection of M's own rows on its PCs."""
        return (self.U[:ncomp] * self.S[:ncomp, None]).t().to(torch.float32).contiguous()

    def cevr(self):
        """Cumulative explained variance ratio (``svd.py:256-261``)."""
        s = self.S.cpu().numpy()
        exp_var = s ** 2 / (s.shape[0] - 1)
        return np.cumsum(exp_var / np.sum(exp_var))


def orthonormalize(Yt, reduce=None, passes=2):
    """Rows of Yt (l,p) -> orthonormal rows spanning the same space (Gram / eigh passes).
    ``reduce``: see :func:`randomized_pcs`."""
    for _ in range(passes):
        G = kernels.cross_gram(Yt, Yt)
        if reduce is not None:
            G = reduce(G)
        Yt = kernels.pcs(_whitening(G), Yt)
    return Yt


def _gram_of_pair(S, l, reduce):
    """Gramian (l,l) fp64 of the rows of ``hi + lo`` from the stacked pair S = [hi; lo] (2l, p)."""
    G2 = kernels.cross_gram(S, S)
    if reduce is not None:
        G2 = reduce(G2)
    G = G2[:l, :l] + G2[:l, l:] + G2[l:, :l] + G2[l:, l:]
    return (0.5 * (G + G.t())).contiguous()


def _whitening(G):
    """Wt with Wt G Wt^T = I (numerically dependent directions zeroed): Cholesky whitening for the small Gramians of
    the sketches (one 150 us launch, no host synchronisation), eigen-based above 128 rows."""
    if G.shape[0] <= 128:
        return kernels.chol_whiten(G)
    evals, evecs, _ = kernels.eigh(G, check=False)
    scale = torch.where(evals > evals[0] * 1e-30, torch.rsqrt(torch.clamp(evals, min=1e-300)), torch.zeros_like(evals))
    return (evecs * scale[:, None]).contiguous()


def orthonormalize_hilo(hi, lo, reduce=None, final=False):
    """Orthonormal basis of the row space of the sketch ``hi + lo`` (error-free fp32 pair, ``kernels.pcs_hilo``).

    The raw sketch is dominated by the leading singular direction -- its rows span (sigma_0/sigma_k)^(2q+1) in
    magnitude -- so its Gramian is formed from the stacked pair in fp64 and the change of basis is applied to the
    pair.  Intermediate bases of the power iteration (``final=False``) are then rounded to plain fp32 (O(1) rows,
    orthonormal to ~1e-4 after the first pass; one ordinary pass finishes): rounding an ITERATE only perturbs a
    self-correcting iteration.  The LAST basis (``final=True``) is returned as a pair (Qhi, Qlo): its span is what the
    principal components inherit, and an fp32 copy would leave 6e-8 of the 1e4-bright halo in every residual."""
    l = hi.shape[0]
    S = torch.cat((hi, lo))                                     # (2l, p)
    Wt = _whitening(_gram_of_pair(S, l, reduce))
    if not final:
        Y1 = kernels.pcs(torch.cat((Wt, Wt), dim=1).contiguous(), S)
        return orthonormalize(Y1, reduce, passes=1)
    l1 = Wt.shape[0]
    S = torch.cat(kernels.pcs_hilo(torch.cat((Wt, Wt), dim=1).contiguous(), S))
    Wt = _whitening(_gram_of_pair(S, l1, reduce))
    return kernels.pcs_hilo(torch.cat((Wt, Wt), dim=1).contiguous(), S)


def randomized_pcs(M, ncomp, omega, n_iter=2, reduce=None, hilo=False, coeffs=False):
    """scikit-learn ``randomized_svd(M, ncomp, n_iter=2, transpose='auto')`` for n < p
    (``svd.py:487-491``): with the Gaussian test matrix ``omega`` (n, ncomp+10) supplied
    by the caller,  Y = M^T (M M^T)^n_iter omega,  Q = orth(Y),  B = Q^T M^T,  PCs = (Q U_B)[:, :k]^T.

    What is computed is the result of that algorithm IN EXACT ARITHMETIC: the sketch is re-orthonormalised after
    every application of M^T M (the row space, hence Q's span and the PCs, is unchanged by a change of basis), the
    raw sketches are carried as error-free fp32 pairs and every reduction runs in fp64.  scikit-learn itself computes
    in the dtype of its input and -- for n_iter = 2 -- without any normalisation, so on a cube that still contains
    the stellar halo (sigma_0/sigma_k ~ 1e3) the reference's fp32 components beyond the first few are rounding noise
    (its residuals differ from one seed to the next by more than their own size) and even its
    float64 run only holds them to ~1e-16 (sigma_0/sigma_k)^5.  On inputs where the reference's arithmetic is sound
    the two agree to 1e-4 with identical ``omega`` (tests/test_gpu_configs.py).

    Pixel-sharded use (``vip_b200/parallel.py``): ``M`` is this rank's column block and
    ``reduce`` sums a small fp64 matrix over the ranks in place (NCCL all-reduce).  Every product
    with the pixel axis contracted -- the (l,n) sketches ``Q M^T`` and the Gramians of the orthonormalisation --
    is a sum over pixel shards; everything else is local.  ``hilo``: return the PCs as the error-free pair
    (Vhi, Vlo) for the high-precision projection.  ``coeffs``: additionally return C (n,k) fp64 = M V^T, the projection
    of M's own rows on the PCs -- it costs nothing: V = Wt Q, so M V^T = (Q M^T)^T Wt^T = B^T Wt^T with the (l,n)
    matrix B the algorithm has already formed (the reference spends a full pass over M on it, pca_fullfr.py:1728)."""
    n, p = M.shape
    red = reduce if reduce is not None else (lambda t: t)
    if isinstance(omega, torch.Tensor):                          # sklearn casts Omega to M's dtype (fp32)
        Om = omega.to(device=M.device, dtype=torch.float32)
    else:
        Om = torch.as_tensor(np.asarray(omega, dtype=np.float32)).to(M.device)
    hi, lo = kernels.pcs_hilo(Om.t().contiguous(), M)            # (l,p) = omega^T M
    for it in range(n_iter):
        Qt = orthonormalize_hilo(hi, lo, reduce)
        Z = red(kernels.cross_gram(Qt, M))                       # (l,n) = Q^T M^T
        hi, lo = kernels.pcs_hilo(Z.contiguous(), M)             # (l,p) = (Q^T M^T) M
    Qhi, Qlo = orthonormalize_hilo(hi, lo, reduce, final=True)
    Qs = torch.cat((Qhi, Qlo))                                   # stacked pair (2l, p)
    l = Qhi.shape[0]
    B2 = red(kernels.cross_gram(Qs, M))                          # (2l,n) fp64
    B = (B2[:l] + B2[l:]).contiguous()                           # (l,n) = Q^T M^T
    evals, evecs, _ = kernels.eigh((B @ B.t()).contiguous(), check=False)
    Wt = evecs[:ncomp].contiguous()                              # rows = leading left vectors of B
    Wt2 = torch.cat((Wt, Wt), dim=1).contiguous()
    V = kernels.pcs_hilo(Wt2, Qs) if hilo else kernels.pcs(Wt2, Qs)
    if coeffs:
        return V, (B.t() @ Wt.t()).contiguous()
    return V


def svd_wrapper(matrix, mode, ncomp, verbose=False, full_output=False, random_state=None, to_numpy=True,
                left_eigv=False):
    """Top-``ncomp`` right singular vectors (ncomp, npix) of a 2-d matrix, computed on the GPU.

    Same call signature as the reference.  ``full_output`` returns, for 'lapack', (left vectors
    (n,ncomp), S, PCs (ncomp,p)) as the reference does; for the other exact modes (U (ncomp,n), S, V)."""
    from .._device import to_device_f32, to_host
    if matrix.ndim != 2:
        raise TypeError("Input matrix is not a 2d array")
    mode = _mode_name(mode)
    is_tensor = isinstance(matrix, torch.Tensor)
    M = matrix.float().contiguous() if is_tensor else to_device_f32(matrix)
    n, p = M.shape
    if ncomp > min(n, p):
        msg = "{} PCs cannot be obtained from a matrix with size [{},{}]."
        msg += " Increase the size of the patches or request less PCs"
        raise RuntimeError(msg.format(ncomp, n, p))
    if mode in _EXACT_MODES:
        if n > p:
            raise NotImplementedError("vip_b200 svd_wrapper expects n_frames <= n_pixels")
        dec = Decomposition(M, None if (full_output or left_eigv) else ncomp)
        V = dec.pcs(ncomp)
        U = dec.U[:ncomp].to(torch.float32)
        S = dec.S[:ncomp].to(torch.float32)
    elif mode in _RAND_MODES:
        if full_output or left_eigv:
            raise NotImplementedError("randsvd with full_output/left_eigv is not implemented")
        rs = random_state
        if rs is None:
            rs = np.random.mtrand._rand          # numpy's global RandomState, like check_random_state(None)
        elif not isinstance(rs, np.random.RandomState):
            rs = np.random.RandomState(rs)
        omega = rs.normal(size=(n, ncomp + 10))
        V = randomized_pcs(M, ncomp, omega)
        U = S = None
    else:
        raise ValueError("The SVD `mode` is not recognized")
    if verbose:
        print("Done SVD/PCA on the GPU (vip_b200, svd_mode={})".format(mode))
    conv = (lambda t: t) if (is_tensor or not to_numpy) else (lambda t: to_host(t))
    if full_output:
        # reference, 'lapack' (svd.py:597-599): (left vectors (n,k), S, PCs (k,p))
        if mode == "lapack":
            return conv(U.t().contiguous()), conv(S), conv(V)
        return conv(U), conv(S), conv(V)
    if left_eigv:
        if mode != "lapack":
            raise NotImplementedError("left_eigv is only implemented for svd_mode='lapack'")
        return conv(U.t().contiguous())
    return conv(V)
