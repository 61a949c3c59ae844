"""The CPU oracle against the UNMODIFIED reference imported from /root/reference (build container
only; skipped on the GPU box where the checkout does not exist)."""
import warnings

import numpy as np
import pytest

from oracle import ref_loader, vip_oracle as O
from tools.synth import adi_cube

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference checkout not present")


@pytest.fixture(scope="module")
def ref():
    warnings.simplefilter("ignore")
    ref_loader.load()
    import vip_hci.psfsub as psfsub
    import vip_hci.preproc as preproc
    return psfsub, preproc


def test_derotate_bit_identical(ref):
    _, preproc = ref
    rng = np.random.default_rng(1)
    angs = np.array([12.3, -33, 77.7, 181, 300.5, 135, 315, 45, -400.2])
    for S in (20, 21):
        cube = rng.normal(size=(9, S, S)).astype(np.float32)
        cube[2, 3, 4] = np.nan
        np.testing.assert_array_equal(preproc.cube_derotate(cube, angs), O.cube_derotate(cube, angs))
        c64 = np.nan_to_num(cube.astype(np.float64))
        np.testing.assert_array_equal(
            preproc.cube_derotate(c64, angs, mask_val=0, interp_zeros=True, ker=1),
            O.cube_derotate(c64, angs, mask_val=0, interp_zeros=True))


def test_pca_and_annular_bit_identical(ref):
    psfsub, _ = ref
    cube, angs = adi_cube(24, 33, 3, 60.0, seed=11)
    for mode in ("lapack", "eigen"):
        r = psfsub.pca(cube, angs, ncomp=3, svd_mode=mode, verbose=False, full_output=True)
        o = O.pca_fullframe(cube, angs, ncomp=3, svd_mode=mode, full_output=True)
        for a, b in zip(r, o):
            np.testing.assert_array_equal(a, b)
    np.random.seed(3)
    r = psfsub.pca(cube, angs, ncomp=3, svd_mode="randsvd", verbose=False)
    np.random.seed(3)
    o = O.pca_fullframe(cube, angs, ncomp=3, svd_mode="randsvd")
    np.testing.assert_array_equal(r, o)
    r = psfsub.pca_annular(cube, angs, ncomp=2, asize=5, verbose=False, full_output=True)
    o = O.pca_annular(cube, angs, ncomp=2, asize=5, full_output=True)
    for a, b in zip(r, o):
        np.testing.assert_array_equal(a, b)


def test_randsvd_restatement(ref):
    cube, _ = adi_cube(24, 33, 3, 60.0, seed=11)
    M = cube.reshape(24, -1)
    om = np.random.RandomState(5).normal(size=(24, 13)).astype(np.float32)
    V1 = O.svd_wrapper(M, "randsvd", 3, random_state=np.random.RandomState(5))
    V2 = O.randsvd_restated(M, 3, om)
    np.testing.assert_allclose(V1.T @ V1, V2.T @ V2, atol=1e-5)


def test_shift_and_median_sub_bit_identical(ref):
    _, preproc = ref
    from vip_hci.psfsub import median_sub
    rng = np.random.default_rng(0)
    for (ny, nx) in ((20, 20), (21, 21), (20, 24), (25, 20)):
        for dt in (np.float32, np.float64):
            fr = rng.normal(size=(ny, nx)).astype(dt)
            for sy, sx in ((1.3, -2.7), (-0.4, 0.2), (3.0, 1.0), (0.0, 0.0), (-5.5, 4.49)):
                a, b = preproc.frame_shift(fr, sy, sx), O.frame_shift(fr, sy, sx)
                assert a.dtype == b.dtype
                np.testing.assert_array_equal(a, b)
    cube, angs = adi_cube(12, 32, 3, 60.0, seed=3)
    for kw in (dict(), dict(collapse="mean")):
        r = median_sub(cube, angs, verbose=False, full_output=True, **kw)
        o = O.median_sub_fullframe(cube, angs, full_output=True, **kw)
        for x, y in zip(r, o):
            np.testing.assert_array_equal(x, y)
    # radius_int > 0 with an ODD number of frames: np.median returns an actual sample, so one residual per pixel
    # is exactly 0 and the reference's default rot_options (mask_val=0, interp_zeros=True; medsub.py:226-229)
    # resets it after the rotation
    cube, angs = adi_cube(11, 32, 3, 60.0, seed=4)
    for kw in (dict(radius_int=3), dict(radius_int=2, collapse="mean"), dict(radius_int=3, mask_val=np.nan)):
        r = median_sub(cube, angs, verbose=False, full_output=True, **kw)
        o = O.median_sub_fullframe(cube, angs, full_output=True, **kw)
        for x, y in zip(r, o):
            np.testing.assert_array_equal(x, y)


def test_pca_incremental_bit_identical(ref):
    """``pca(..., batch=...)`` (incremental PCA, utils_pca.py:431-614): frame, PCs and per-batch frames."""
    psfsub, _ = ref
    cube, angs = adi_cube(23, 33, 3, 60.0, seed=11)
    for batch in (6, 23, 10):
        r = psfsub.pca(cube, angs, ncomp=3, batch=batch, verbose=False, full_output=True)
        o = O.pca_incremental(cube, angs, batch, ncomp=3, full_output=True)
        assert len(r) == len(o) == 3
        for a, b in zip(r, o):
            np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(psfsub.pca(cube, angs, ncomp=2, batch=8, collapse="mean", verbose=False),
                                  O.pca_incremental(cube, angs, 8, ncomp=2, collapse="mean"))
    with pytest.raises(ValueError):
        O.pca_incremental(cube, angs, 2, ncomp=3)          # first batch smaller than ncomp (scikit-learn's check)


def test_snr_and_snrmap_logic_bit_identical(ref):
    """``metrics.snr`` / ``snrmap`` of the unmodified reference (with the photutils stand-in of oracle/ref_loader.py
    supplying the exact aperture sums) against the oracle: aperture centres, small-sample statistics, the
    exclude_negative_lobes / array2 / use2alone options and the pixel mask of the map."""
    from vip_hci.metrics.snr_source import snr, snrmap, indep_ap_centers
    rng = np.random.default_rng(5)
    a = rng.normal(size=(40, 41))
    a2 = rng.normal(size=(40, 41))
    a[3, 7] = 0.0                                   # a zero pixel drops out of the map
    for xy in ((30, 22), (12.5, 9.25), (33, 5)):
        for kw in (dict(), dict(exclude_negative_lobes=True), dict(array2=a2), dict(array2=a2, use2alone=True)):
            r = snr(a, xy, 4.0, full_output=True, **kw)
            o = O.snr(a, xy, 4.0, full_output=True, **kw)
            for x, y in zip(r, o):
                np.testing.assert_array_equal(x, y)
        np.testing.assert_array_equal(np.array(indep_ap_centers(a, xy, 4.0, exclude_theta_range=(20, 70))),
                                      np.array(O.indep_ap_centers(a, xy, 4.0, exclude_theta_range=(20, 70))))
    with pytest.raises(RuntimeError):
        O.snr(a, (20.5, 20.2), 4.0)
    m_ref = snrmap(a, 4.0, nproc=1, verbose=False)
    np.testing.assert_array_equal(m_ref, O.snrmap(a, 4.0))
    m_ref = snrmap(a, 3.0, nproc=1, verbose=False, array2=a2, exclude_negative_lobes=True)
    np.testing.assert_array_equal(m_ref, O.snrmap(a, 3.0, array2=a2, exclude_negative_lobes=True))


def test_pca_annular_adimsdi_bit_identical(ref):
    """``pca_annular(cube4d, ..., scale_list=, ncomp=(k_ifs, k_adi))`` (``pca_local.py:332-462``, ``_pca_sdi_fr``
    :470-591) of the unmodified reference against the oracle: both passes, and the single-pass variant (k_adi None)."""
    psfsub, _ = ref
    from tools.make_golden import ifs_cube
    cube, angs, sl = ifs_cube(z=5, n=8, size=24, seed=7)
    for ncomp, kw in (((2, 2), dict(asize=4, delta_sep=(0.1, 0.3))), ((1, None), dict(asize=6, delta_sep=0.1)),
                      ((2, 2), dict(asize=4, delta_sep=(0.05, 0.15), n_segments=2, collapse_ifs="median",
                                    scaling="temp-mean"))):
        r = psfsub.pca_annular(cube, angs, scale_list=sl, ncomp=ncomp, fwhm=3, verbose=False, full_output=True,
                               nproc=1, **kw)
        o = O.pca_annular_sdi(cube, angs, sl, ncomp, fwhm=3, full_output=True, **kw)
        assert len(r) == len(o) == 3
        for a, b in zip(r, o):
            np.testing.assert_array_equal(a, b)
    # reference cube: the same spectral pass on `cube_ref`, its residual frames as the library of the ADI pass
    cref, _, _ = ifs_cube(z=5, n=6, size=24, seed=8)
    r = psfsub.pca_annular(cube, angs, scale_list=sl, ncomp=(2, 2), fwhm=3, asize=4, delta_sep=(0.1, 0.3),
                           cube_ref=cref, verbose=False, full_output=True, nproc=1)
    o = O.pca_annular_sdi(cube, angs, sl, (2, 2), fwhm=3, asize=4, delta_sep=(0.1, 0.3), cube_ref=cref,
                          full_output=True)
    for a, b in zip(r, o):
        np.testing.assert_array_equal(a, b)


def test_pca_annular_ncomp_auto_bit_identical(ref):
    """``pca_annular(ncomp='auto', tol=)``: the noise-decay rule of ``get_eigenvectors`` (``psfsub/svd.py:622-672``)
    through ``do_pca_patch``, for two tolerances (different numbers of components per patch) and a tuple mixing
    'auto' with fixed numbers."""
    psfsub, _ = ref
    cube, angs = adi_cube(24, 40, 4, 80.0, seed=5)
    for kw in (dict(ncomp="auto", tol=0.1, asize=5, delta_rot=(0.1, 0.4)),
               dict(ncomp="auto", tol=0.5, asize=6, n_segments=2, delta_rot=0.3),
               dict(ncomp="auto", tol=0.02, asize=5, delta_rot=0.2)):
        r = psfsub.pca_annular(cube, angs, verbose=False, full_output=True, nproc=1, **kw)
        used = []
        o = O.pca_annular(cube, angs, full_output=True, ncomp_out=used, **kw)
        for a, b in zip(r, o):
            np.testing.assert_array_equal(a, b)
        assert len(set(used)) > 1 or kw["tol"] == 0.5, (kw, sorted(set(used)))


def test_pca_annular_left_eigv_bit_identical(ref):
    """``pca_annular(left_eigv=True)``: projection of every segment on the temporal singular vectors of the pixels
    outside it (``pca_local.py:704-707, 755-779``)."""
    psfsub, _ = ref
    cube, angs = adi_cube(16, 32, 3, 70.0, seed=8)
    for kw in (dict(ncomp=2, asize=5), dict(ncomp=3, asize=4, n_segments=2, scaling="temp-mean"),
               dict(ncomp=2, asize=5, scaling="spat-mean")):
        r = psfsub.pca_annular(cube, angs, left_eigv=True, verbose=False, full_output=True, nproc=1, **kw)
        o = O.pca_annular(cube, angs, left_eigv=True, full_output=True, **kw)
        for a, b in zip(r, o):
            np.testing.assert_array_equal(a, b)


def _ifs_small():
    from tools.make_golden import ifs_cube
    cube, angs, sl = ifs_cube(z=5, n=10, size=24, seed=43)
    cref = ifs_cube(z=5, n=6, size=24, seed=47)[0]
    return cube, angs, sl, cref


def test_pca_adimsdi_fullframe_options_bit_identical(ref):
    """Full-frame ADI+mSDI options of round 2 (second half): reference cube in both modes (``pca_fullfr.py:499-510,
    1099-1119, 1278-1282, 1392-1404``), PA-rejection second pass with ``source_xy`` and ``cube_sig`` (:1407-1460), and
    the single-pass PCA grid with / without ``source_xy`` (:1205-1236 -> ``utils_pca.py:191-228``)."""
    psfsub, _ = ref
    cube, angs, sl, cref = _ifs_small()
    kw = dict(scale_list=sl, verbose=False)
    # double pass, RSDI
    r = psfsub.pca(cube, angs, cube_ref=cref, adimsdi="double", ncomp=(2, 3), full_output=True, **kw)
    o = O.pca_adimsdi_double(cube, angs, sl, (2, 3), cube_ref=cref, full_output=True)
    for a, b in zip(r, o):
        np.testing.assert_array_equal(a, b)
    # double pass, source_xy (+ reference cube, + cube_sig)
    sig = np.zeros(cube.shape[1:], dtype=np.float32)
    sig[:, 12, 17] = 3.0
    for extra in (dict(), dict(cube_ref=cref), dict(cube_sig=sig)):
        r = psfsub.pca(cube, angs, adimsdi="double", ncomp=(2, 2), source_xy=(17, 12), delta_rot=0.5, fwhm=3,
                       min_frames_pca=2, full_output=True, **kw, **extra)
        o = O.pca_adimsdi_double(cube, angs, sl, (2, 2), source_xy=(17, 12), delta_rot=0.5, fwhm=3, min_frames_pca=2,
                                 full_output=True, **extra)
        for a, b in zip(r, o):
            np.testing.assert_array_equal(a, b)
    # single pass with a reference cube, both strategies
    for strat, tag in (("RDI", "RSDI"), ("ARDI", "ARSDI")):
        r = psfsub.pca(cube, angs, cube_ref=cref, adimsdi="single", ncomp=3, ref_strategy=strat, full_output=True, **kw)
        o = O.pca_adimsdi_single(cube, angs, sl, 3, cube_ref=cref, ref_strategy=tag, full_output=True)
        for a, b in zip(r, o):
            np.testing.assert_array_equal(a, b)
    # single-pass grid
    r, pcl = psfsub.pca(cube, angs, adimsdi="single", ncomp=(1, 3), full_output=True, **kw)
    o, opcl = O.pca_adimsdi_single_grid(cube, angs, sl, (1, 3))
    assert list(pcl) == list(opcl)
    np.testing.assert_array_equal(r, o)
    r = psfsub.pca(cube, angs, adimsdi="single", ncomp=[2, 4], ifs_collapse_range=(1, 4), collapse="mean", **kw)
    o, _ = O.pca_adimsdi_single_grid(cube, angs, sl, [2, 4], ifs_collapse_range=(1, 4), collapse="mean")
    np.testing.assert_array_equal(r, o)
    r = psfsub.pca(cube, angs, adimsdi="single", ncomp=(1, 3), source_xy=(17, 12), fwhm=3, full_output=True, **kw)
    o = O.pca_adimsdi_single_grid(cube, angs, sl, (1, 3), source_xy=(17, 12), fwhm=3)
    np.testing.assert_array_equal(r[0], o[0])
    np.testing.assert_array_equal(r[1], o[1])
    # the reference averages an OBJECT array (sequential Python sums), the oracle a float array (pairwise): 1 ulp
    np.testing.assert_allclose(np.asarray(r[2]["S/Ns"], dtype=float), np.asarray(o[2]["S/Ns"], dtype=float), rtol=1e-13)
    np.testing.assert_allclose(np.asarray(r[2]["fluxes"], dtype=float), np.asarray(o[2]["fluxes"], dtype=float),
                               rtol=1e-13)


def _rdi_masks(size):
    from oracle import vip_oracle as O_
    ones = np.ones((size, size))
    yy, xx = O_.get_annulus_segments((size, size), 3, 7)[0]
    boat = np.zeros((size, size)); boat[yy, xx] = 1
    yy, xx = O_.get_annulus_segments((size, size), 8, 6)[0]
    anchor = np.zeros((size, size)); anchor[yy, xx] = 1
    return anchor, boat


def test_pca_mask_rdi_bit_identical(ref):
    """``pca(cube, angs, cube_ref=ref, mask_rdi=(anchor, boat), ncomp=k)``: PCA with data imputation
    (``pca_fullfr.py:966-972`` -> ``cube_subtract_sky_pca``, ``preproc/skysubtraction.py:36-260``), two masks and one."""
    psfsub, _ = ref
    cube, angs = adi_cube(16, 33, 3, 60.0, seed=21)
    cref = adi_cube(12, 33, 3, 60.0, seed=22)[0]
    anchor, boat = _rdi_masks(33)
    for masks in ((anchor, boat), anchor):
        for dt in (np.float32, np.float64):
            r = psfsub.pca(cube.astype(dt), angs, cube_ref=cref.astype(dt), mask_rdi=masks, ncomp=3, verbose=False,
                           full_output=True)
            o = O.pca_fullframe(cube.astype(dt), angs, cube_ref=cref.astype(dt), mask_rdi=masks, ncomp=3,
                                full_output=True)
            assert len(r) == len(o) == 5
            for a, b in zip(r, o):
                assert a.dtype == b.dtype and a.shape == b.shape
                np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(psfsub.pca(cube, angs, cube_ref=cref, mask_rdi=(anchor, boat), ncomp=2, verbose=False),
                                  O.pca_fullframe(cube, angs, cube_ref=cref, mask_rdi=(anchor, boat), ncomp=2))
