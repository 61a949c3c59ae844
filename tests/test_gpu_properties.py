"""Size-independent properties of the CUDA path AT BASELINE SIZES (the oracle needs minutes to hours there):
linearity and frame independence of the vip-fft derotation (512^2 planes of config 2, 1024^2 planes of config 5),
the projector identities of the PCA stage and exact scale / permutation equivariance of the whole config-2 call, and the
order statistics identities of the median collapse.  Parity with the reference at these sizes is covered by the golden
fixtures of tests/test_gpu_configs.py; these tests need no reference at all."""
import numpy as np
import pytest

from tools.synth import adi_cube
from conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vb():
    import vip_b200
    return vip_b200


@pytest.fixture(scope="module")
def c2():
    return adi_cube(500, 512, 20, 90.0, seed=20260102)


@pytest.mark.parametrize("size", [512, 1024])
def test_derotation_is_linear_and_frame_independent(vb, size):
    """``cube_derotate`` is a linear map per frame (three FFT shears, ``preproc/derotation.py:542-640``):
    R(2X - 3Y) = 2 R(X) - 3 R(Y) to fp32 rounding, and frame i of a batch equals the same frame rotated alone,
    bit for bit (the kernels never mix frames)."""
    rng = np.random.default_rng(size)
    n = 6 if size == 512 else 3
    X = rng.normal(size=(n, size, size)).astype(np.float32)
    Y = rng.normal(size=(n, size, size)).astype(np.float32)
    angs = np.array([7.3, -58.1, 133.0, 201.7, 310.2, 89.9])[:n]
    rx, ry = vb.cube_derotate(X, angs), vb.cube_derotate(Y, angs)
    rz = vb.cube_derotate(2.0 * X - 3.0 * Y, angs)
    assert rel_err(rz, 2.0 * rx.astype(np.float64) - 3.0 * ry) < 2e-6
    # exact power-of-two scaling and exact negation
    assert np.array_equal(vb.cube_derotate(4.0 * X, angs), 4.0 * rx, equal_nan=True)
    assert np.array_equal(vb.cube_derotate(-X, angs), -rx, equal_nan=True)
    for i in (0, n - 1):
        alone = vb.cube_derotate(X[i:i + 1].copy(), angs[i:i + 1])
        assert np.array_equal(alone[0], rx[i], equal_nan=True)


def test_c2_projector_identities(vb, c2):
    """PCA stage of config 2 at full size (``_project_subtract``, ``pca_fullfr.py:1552-1737``): the PCs are
    orthonormal, reconstruction + residuals give the cube back, and the residuals carry nothing along the PCs
    (idempotence of I - V^T V)."""
    cube, angs = c2
    frame, pcs, recon, res, res_ = vb.pca(cube, angs, ncomp=20, verbose=False, full_output=True)
    V = pcs.reshape(20, -1).astype(np.float64)
    # the Gramian's products are fp32-grade (error-free bf16x3 split): lambda_k carries ~6e-8 lambda_0 / lambda_k
    assert np.max(np.abs(V @ V.T - np.eye(20))) < 1e-5
    sel = [0, 123, 499]
    M = cube[sel].reshape(3, -1).astype(np.float64)
    R = res[sel].reshape(3, -1).astype(np.float64)
    assert np.max(np.abs(R + recon[sel].reshape(3, -1) - M)) < 2e-6 * np.max(np.abs(M))
    coeff = M @ V.T
    assert np.max(np.abs(R @ V.T)) < 1e-5 * np.max(np.abs(coeff))
    assert frame.shape == (512, 512) and np.isfinite(frame).all()


def test_c2_scale_and_permutation_equivariance(vb, c2):
    """The whole config-2 call commutes with an exact rescaling of the data (a power of two: every fp32 / bf16 / fp64
    intermediate scales exactly, only absolute thresholds could break it) and with a permutation of the frames
    (summation orders change: fp32 rounding)."""
    cube, angs = c2
    frame = vb.pca(cube, angs, ncomp=20, verbose=False)
    f4 = vb.pca(4.0 * cube, angs, ncomp=20, verbose=False)
    # not bit-exact: the eigensolver accumulates its small Gramians with atomics, so two runs differ by fp32 ulps of
    # the residuals (~3e-5 near the star) -- a few 1e-7 of the frame peak; passed at 1e-6 on the B200, bound 5e-6
    assert rel_err(f4, 4.0 * frame.astype(np.float64)) < 5e-6
    perm = np.random.default_rng(3).permutation(cube.shape[0])
    fp = vb.pca(np.ascontiguousarray(cube[perm]), angs[perm], ncomp=20, verbose=False)
    assert rel_err(fp, frame) < 2e-5


def test_c2_median_order_statistics(vb, c2):
    """Median collapse over 500 frames of 512^2 (``cube_collapse``, ``preproc/subsampling.py:79-112``): invariant
    under a permutation of the frames, odd under negation, exact under power-of-two scaling -- bit for bit -- and
    bracketed by the 249th / 250th order statistics computed by numpy on a pixel subset."""
    cube, _ = c2
    med = vb.cube_collapse(cube, mode="median")
    perm = np.random.default_rng(4).permutation(cube.shape[0])
    assert np.array_equal(vb.cube_collapse(np.ascontiguousarray(cube[perm]), mode="median"), med)
    assert np.array_equal(vb.cube_collapse(-cube, mode="median"), -med)
    assert np.array_equal(vb.cube_collapse(0.5 * cube, mode="median"), 0.5 * med)
    sub = np.sort(cube[:, ::37, ::41], axis=0)
    want = (0.5 * (sub[249].astype(np.float64) + sub[250])).astype(np.float32)     # np.median of an even count
    assert np.array_equal(med[::37, ::41], want)
    odd = vb.cube_collapse(cube[:499], mode="median")
    assert np.array_equal(odd[::37, ::41], np.sort(cube[:499, ::37, ::41], axis=0)[249])
