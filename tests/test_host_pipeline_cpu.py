"""Host orchestration of the drop-in callables on CPU: ``vip_b200.pca`` / ``median_sub`` / ``cube_derotate`` /
``cube_collapse`` run end to end with the kernels replaced by the CPU stand-ins of ``tests/kernel_double.py`` and
are compared with the golden outputs of the unmodified reference and with the oracle.  This pins everything that
is NOT a kernel -- parameter parsing, branch selection, scaling / masking / library assembly, return layouts and
dtypes -- in the container that has no GPU; the kernels themselves are pinned on the GPU (test_gpu_parity.py).
"""
import numpy as np
import pytest

import kernel_double
from oracle import vip_oracle as O
from tools.synth import adi_cube
from conftest import rel_err

TOL = 2e-5       # stand-ins are fp64-accurate; the slack is the reference's own fp32 arithmetic


@pytest.fixture
def vb(monkeypatch):
    return kernel_double.install(monkeypatch)


def test_pca_c1_layouts_and_values(vb, golden, golden_inputs):
    g = golden["pca_fullframe"]
    cube, angs = golden_inputs["c1"]
    fr, pcs, recon, res, res_ = vb.pca(cube, angs, ncomp=5, verbose=False, full_output=True)
    assert fr.dtype == np.float32 and fr.shape == (101, 101)
    assert pcs.shape == (5, 101, 101) and recon.shape == res.shape == res_.shape == cube.shape
    scale = np.max(np.abs(g["c1_res_frame7"]))
    assert np.max(np.abs(res[7] - g["c1_res_frame7"])) < 1e-4 * scale
    assert np.max(np.abs(res_[7] - g["c1_resder_frame7"])) < 1e-4 * scale
    assert rel_err(fr, g["c1_frame"]) < 3e-4
    P = pcs.reshape(5, -1)
    assert np.max(np.abs((P.T @ P)[::97, ::89] - g["c1_proj"])) < 1e-5
    assert rel_err(vb.pca(cube, angs, ncomp=5, verbose=False), g["c1_frame"]) < 3e-4


def test_pca_options(vb, golden, golden_inputs):
    g = golden["pca_fullframe"]
    cube, angs = golden_inputs["small"]
    for mode in ("lapack", "eigen", "arpack"):
        assert rel_err(vb.pca(cube, angs, ncomp=4, svd_mode=mode, verbose=False), g["small_lapack"]) < 3e-4
    for sc in ("temp-mean", "spat-mean", "temp-standard", "spat-standard"):
        assert rel_err(vb.pca(cube, angs, ncomp=3, scaling=sc, verbose=False), g[f"small_{sc}"]) < 3e-4, sc
    for col in ("mean", "sum"):
        assert rel_err(vb.pca(cube, angs, ncomp=3, collapse=col, verbose=False), g[f"small_{col}"]) < 3e-4
    ref = adi_cube(20, 41, 4, 60.0, seed=6)[0]
    c64, r64 = cube.astype(np.float64), ref.astype(np.float64)
    assert rel_err(vb.pca(cube, angs, cube_ref=ref, ncomp=4, verbose=False),
                   O.pca_fullframe(c64, angs, ncomp=4, cube_ref=r64)) < 1e-4
    assert rel_err(vb.pca(cube, angs, cube_ref=ref, ncomp=4, ref_strategy="ARDI", verbose=False),
                   O.pca_fullframe(c64, angs, ncomp=4, cube_ref=np.concatenate((c64, r64)))) < 1e-4
    assert rel_err(vb.pca(cube, angs, ncomp=0.9995, verbose=False), g["small_cevr"]) < 3e-4
    fr = vb.pca(cube, angs, None, None, 4, "lapack", verbose=False)          # positional, dataclass order
    assert rel_err(fr, g["small_lapack"]) < 3e-4
    from vip_b200.psfsub import PCA_Params
    fr = vb.pca(algo_params=PCA_Params(cube=cube, angle_list=angs, ncomp=4, verbose=False))
    assert rel_err(fr, g["small_lapack"]) < 3e-4


def test_pca_mask_sig_grid_4d(vb, golden, golden_inputs):
    cube, angs = golden_inputs["small"]
    assert rel_err(vb.pca(cube, angs, ncomp=3, mask_center_px=4, verbose=False),
                   O.pca_fullframe(cube, angs, ncomp=3, mask_center_px=4)) < 3e-4
    sig = np.zeros_like(cube)
    sig[:, 30:34, 20:24] = 5.0
    assert rel_err(vb.pca(cube, angs, ncomp=3, cube_sig=sig, verbose=False),
                   O.pca_fullframe(cube, angs, ncomp=3, cube_sig=sig)) < 3e-4
    g = golden["pca_grid4d"]
    fr, pcl = vb.pca(cube, angs, ncomp=(1, 4), verbose=False, full_output=True)
    assert pcl == list(g["grid_range_pclist"]) and fr.dtype == np.float32
    for i in range(len(pcl)):
        assert rel_err(fr[i], g["grid_range"][i]) < 3e-4, i
    assert rel_err(vb.pca(cube, angs, ncomp=[2, 4], verbose=False), g["grid_list"]) < 3e-4
    assert rel_err(vb.pca(cube, angs, ncomp=(1, 5, 2), med_of_npcs=True, verbose=False), g["grid_step_med"]) < 3e-4
    cube4, angs4, _ = golden_inputs["ifs"]
    r = vb.pca(cube4, angs4, ncomp=2, verbose=False, full_output=True)
    assert len(r) == 6 and r[0].dtype == np.float64 and r[5].dtype == np.float64 and r[1].dtype == np.float32
    assert rel_err(r[0], g["ch_frame"]) < 3e-4 and rel_err(r[5], g["ch_ifs"]) < 3e-4
    final, pcl4, ifs = vb.pca(cube4, angs4, ncomp=[1, 3], verbose=False, full_output=True)
    assert pcl4 == [[1, 3]] * 6
    assert rel_err(final, g["ch_grid"]) < 3e-4 and rel_err(ifs, g["ch_grid_ifs"]) < 3e-4


def test_pca_errors_and_clamp(vb):
    cube, angs = adi_cube(8, 16, 2, 30.0, seed=1)
    with pytest.raises(ValueError):
        vb.pca(cube, angs[:-1], ncomp=2, verbose=False)
    with pytest.raises(ValueError):
        vb.pca(cube, angs, ncomp=0, verbose=False)
    assert vb.pca(cube, angs, ncomp=50, verbose=False).shape == (16, 16)


def test_derotate_collapse_median_sub(vb, golden, golden_inputs):
    cube, angs = golden_inputs["derot33"]
    out = vb.cube_derotate(cube, angs)
    assert out.dtype == cube.dtype and rel_err(out, golden["derotate"]["derot33"]) < TOL
    small = golden_inputs["small"][0].copy()
    small[3, 5, 5] = np.nan
    small[:, 7, 7] = np.nan
    for m in ("median", "mean", "sum", "max", "absmean"):
        np.testing.assert_array_equal(vb.cube_collapse(small, m), golden["collapse"][m], err_msg=m)
    cube, angs = golden_inputs["small"]
    co, cd, fr = vb.median_sub(cube, angs, full_output=True, verbose=False)
    oo, od, of = O.median_sub_fullframe(cube, angs, full_output=True)
    np.testing.assert_array_equal(co, oo)
    assert rel_err(cd, od) < TOL and rel_err(fr, of) < TOL


def test_pca_incremental_vs_oracle(vb):
    """``pca(..., batch=...)``: the mini-batch model (mean update, stacked-matrix SVD through the Gramian route,
    sign convention), the second pass and the median of the batch frames, against the oracle (= scikit-learn's
    IncrementalPCA, bit-identical to the reference: tests/test_oracle_vs_reference.py)."""
    cube, angs = adi_cube(23, 33, 3, 60.0, seed=11)
    for batch in (6, 10, 23):
        fr, pcs, med = vb.pca(cube, angs, ncomp=3, batch=batch, verbose=False, full_output=True)
        ofr, opcs, omed = O.pca_incremental(cube, angs, batch, ncomp=3, full_output=True)
        assert fr.dtype == pcs.dtype == med.dtype == np.float64
        assert fr.shape == ofr.shape and pcs.shape == opcs.shape and med.shape == omed.shape
        assert rel_err(med, omed) < 1e-4, batch
        assert rel_err(fr, ofr) < 3e-4, batch          # frame ~1e-3 of the batch frames: fp32 residuals vs float64
        assert np.max(np.abs(pcs - opcs)) < 1e-4 * np.max(np.abs(opcs)), batch     # same signs (svd_flip rule)
    assert rel_err(vb.pca(cube, angs, ncomp=2, batch=8, collapse="mean", verbose=False),
                   O.pca_incremental(cube, angs, 8, ncomp=2, collapse="mean")) < 1e-4
    assert vb.pca(cube, angs, ncomp=2, batch=0.5, verbose=False).shape == (33, 33)     # memory-sized: one batch
    res = vb.psfsub.pca_incremental(cube, angs, batch=7, ncomp=3, verbose=False, return_residuals=True)
    assert rel_err(res, O.pca_incremental(cube, angs, 7, ncomp=3, return_residuals=True)) < 1e-4
    with pytest.raises(ValueError):
        vb.pca(cube, angs, ncomp=3, batch=2, verbose=False)            # first batch smaller than ncomp
    with pytest.raises(ValueError):
        vb.pca(cube, angs, cube_ref=cube, ncomp=3, batch=6, verbose=False)
    with pytest.raises(TypeError):
        vb.pca(cube, angs, ncomp=3, batch="6", verbose=False)


def test_pca_incremental_golden(vb, golden, golden_inputs):
    from tools.make_golden import INCREMENTAL_CASES
    g = golden["pca_incremental"]
    cube, angs = golden_inputs["small"]
    for key, kw in INCREMENTAL_CASES.items():
        fr, pcs, med = vb.pca(cube, angs, verbose=False, full_output=True, **kw)
        assert rel_err(med, g[f"{key}_medians"]) < 1e-4 and rel_err(fr, g[f"{key}_frame"]) < 3e-4, key
        assert np.max(np.abs(pcs - g[f"{key}_pcs"])) < 1e-4 * np.max(np.abs(g[f"{key}_pcs"])), key


def test_pca_annular_layouts_and_values(vb, golden, golden_inputs):
    g = golden["pca_annular"]
    cube, angs = golden_inputs["ann"]
    co, cd, fr = vb.pca_annular(cube, angs, ncomp=3, asize=6, verbose=False, full_output=True)
    assert co.shape == cd.shape == cube.shape and co.dtype == np.float32
    scale = np.max(np.abs(g["ann_cube_out5"]))
    assert np.max(np.abs(co[5] - g["ann_cube_out5"])) < 1e-4 * scale
    assert rel_err(fr, g["ann_frame"]) < 3e-4
    fr = vb.pca_annular(cube, angs, ncomp=2, asize=6, n_segments=3, delta_rot=0.5, radius_int=4, verbose=False)
    assert rel_err(fr, g["ann_seg_frame"]) < 3e-4
    co4, cd4, frl = vb.pca_annular(cube, angs, ncomp=[1, 3], asize=6, verbose=False, full_output=True)
    assert co4.shape == (2,) + cube.shape and co4.dtype == np.float64 and isinstance(frl, list) and len(frl) == 2
    for i in range(2):
        assert rel_err(frl[i], g["ann_list_frames"][i]) < 3e-4
    ref = adi_cube(12, 48, 3, 80.0, seed=10)[0]
    sig = np.zeros_like(cube)
    sig[:, 30:33, 10:13] = 4.0
    for kw in (dict(ncomp=(1, 2, 3, 2), asize=6), dict(ncomp=2, asize=6, cube_sig=sig), dict(ncomp=2, asize=6, cube_ref=ref),
               dict(ncomp=2, asize=6, scaling="temp-mean"), dict(ncomp=2, asize=8, max_frames_lib=12)):
        o = O.pca_annular(cube, angs, full_output=True, **kw)
        r = vb.pca_annular(cube, angs, full_output=True, verbose=False, **kw)
        assert np.max(np.abs(r[0] - o[0])) < 1e-4 * np.max(np.abs(o[0])), kw
        assert rel_err(r[2], o[2]) < 3e-4, kw
    with pytest.raises(TypeError):
        vb.pca_annular(cube, angs[:-1], ncomp=2, asize=6, verbose=False)
    with pytest.raises(RuntimeError):
        vb.pca_annular(cube, angs, ncomp=2, asize=6, delta_rot=500, verbose=False)
    g4 = golden["pca_annular_4d"]
    cube4, angs4, _ = golden_inputs["ifs"]
    co, cd, fr = vb.pca_annular(cube4[:3], angs4, ncomp=2, asize=5, delta_rot=(0.05, 0.2), verbose=False, full_output=True)
    assert co.dtype == np.float32 and fr.dtype == np.float64 and rel_err(fr, g4["ann4d_frame"]) < 3e-4


def test_pca_sdi_layouts_and_values(vb, golden, golden_inputs):
    g = golden["pca_sdi"]
    cube, angs, sl = golden_inputs["ifs"]
    tol = 5e-6 * float(np.max(np.abs(cube)))
    fr, rc, rd = vb.pca(cube, angs, scale_list=sl, adimsdi="double", ncomp=(2, 3), verbose=False, full_output=True)
    assert fr.dtype == np.float64 and rc.shape == (cube.shape[1],) + cube.shape[2:]
    assert np.max(np.abs(rc - g["double_res_channels"])) < tol and np.max(np.abs(fr - g["double_frame"])) < tol
    fr = vb.pca(cube, angs, scale_list=sl, adimsdi="double", ncomp=(2, None), verbose=False)
    assert np.max(np.abs(fr - g["double_skipadi"])) < tol
    with pytest.raises(TypeError):
        vb.pca(cube, angs, scale_list=sl, adimsdi="double", ncomp=3, verbose=False)
    gs = golden["pca_sdi_single"]
    fr, allfr, desc, resadi = vb.pca(cube, angs, scale_list=sl, adimsdi="single", ncomp=3, verbose=False, full_output=True)
    assert fr.dtype == np.float64 and allfr.shape == (84, 32, 32) and desc.shape == cube.shape
    assert np.max(np.abs(desc[2] - gs["single_desc_ch2"])) < tol and np.max(np.abs(fr - gs["single_frame"])) < tol


def test_pca_source_xy_and_left_eigv(vb, golden, golden_inputs):
    """Frame-by-frame PCA with PA-rejection libraries (one Gramian, per-frame sub-problems) and ``left_eigv``."""
    from tools.make_golden import SOURCE_XY_CASES
    g = golden["pca_source_xy"]
    cube, angs = golden_inputs["small"]
    ref = adi_cube(20, 41, 4, 60.0, seed=6)[0]
    for key, kw in SOURCE_XY_CASES.items():
        extra = dict(cube_ref=ref) if key == "rdi" else {}
        fr, recon, res, res_ = vb.pca(cube, angs, verbose=False, full_output=True, **kw, **extra)
        assert fr.dtype == np.float32 and recon.shape == cube.shape and res_.shape == cube.shape
        assert rel_err(res, g[f"{key}_res"]) < 1e-4, key
        c64 = cube.astype(np.float64)
        e64 = {k: v.astype(np.float64) for k, v in extra.items()}
        truth = O.pca_fullframe(c64, angs, **kw, **e64)
        assert rel_err(fr, g[f"{key}_frame"]) < 3e-4 or rel_err(fr, truth) < 1.5 * rel_err(g[f"{key}_frame"], truth) + 2e-5
    gl = golden["pca_left_eigv"]
    fr, pcs, recon, res, _ = vb.pca(cube, angs, ncomp=4, left_eigv=True, verbose=False, full_output=True)
    assert pcs.shape == gl["left_pcs"].shape
    assert rel_err(res, gl["left_res"]) < 1e-4 and rel_err(fr, gl["left_frame"]) < 3e-4


def test_pca_check_memory(vb, monkeypatch):
    """``check_memory`` (pca_fullfr.py:438-455): an input larger than the available (device) memory raises
    RuntimeError pointing at ``batch``; ``check_memory=False`` and ``batch`` bypass the check."""
    from vip_b200 import _device
    from vip_b200.psfsub import pca_fullfr
    cube, angs = adi_cube(12, 16, 2, 40.0, seed=5)
    monkeypatch.setattr(pca_fullfr, "_MEMCHECK_MIN_BYTES", 0)      # the query is skipped for inputs below 8 GiB
    monkeypatch.setattr(_device, "free_memory_bytes", lambda: cube.nbytes - 1)
    with pytest.raises(RuntimeError, match="batch"):
        vb.pca(cube, angs, ncomp=2, verbose=False)
    assert vb.pca(cube, angs, ncomp=2, verbose=False, check_memory=False).shape == (16, 16)
    assert vb.pca(cube, angs, ncomp=2, batch=6, verbose=False).shape == (16, 16)


def test_randsvd_is_the_exact_arithmetic_result_of_sklearns_algorithm(vb, golden_inputs):
    """``svd_mode='randsvd'`` (psfsub/svd.py randomized_pcs) through the stand-ins: with identical Omega the
    residuals equal those of scikit-learn's ``randomized_svd`` evaluated in float64 (where its unnormalised power
    iterations are still accurate on this cube), for a halo-dominated fp32 cube on which the reference's OWN fp32
    run is rounding noise (checked here too: O(1) away from its float64 run)."""
    cube, angs = golden_inputs["small"]
    import torch
    from vip_b200.psfsub.pca_fullfr import project_subtract_device
    res = project_subtract_device(torch.from_numpy(cube), 4, svd_mode="randsvd",
                                  random_state=np.random.RandomState(11)).numpy()
    o64 = O.project_subtract(cube.astype(np.float64), 4, svd_mode="randsvd", random_state=np.random.RandomState(11))
    o32 = O.project_subtract(cube, 4, svd_mode="randsvd", random_state=np.random.RandomState(11))
    scale = np.max(np.abs(o64))
    assert np.max(np.abs(res - o64)) / scale < 1e-5
    assert np.max(np.abs(o32 - o64)) / scale > 0.1            # the reference's fp32 arithmetic has lost the subspace
    # a spectrum gapped at ncomp: the randomized subspace is the exact one
    exact = O.project_subtract(cube.astype(np.float64), 4, svd_mode="lapack")
    assert np.max(np.abs(res - exact)) / scale < 1e-5
    # end to end (global RandomState like the reference), frame against the float64 reference run
    np.random.seed(3)
    fr = vb.pca(cube, angs, ncomp=4, svd_mode="randsvd", verbose=False)
    np.random.seed(3)
    assert rel_err(fr, O.pca_fullframe(cube.astype(np.float64), angs, ncomp=4, svd_mode="randsvd")) < 1e-4


def test_snr_snrmap_and_snr_optimised_grid_host_logic(vb):
    """Host side of ``vip_b200.snr`` / ``snrmap`` and of ``pca(source_xy=, ncomp=(lo, hi))`` (return layouts, masks,
    option handling) through the stand-ins, against the oracle."""
    rng = np.random.default_rng(5)
    a = rng.normal(size=(40, 41)).astype(np.float32)
    a[5, 20] = 0.0
    for kw in (dict(), dict(exclude_negative_lobes=True), dict(exclude_theta_range=(20, 70))):
        r = vb.snr(a, (30, 22), 4.0, full_output=True, **kw)
        o = O.snr(a, (30, 22), 4.0, full_output=True, **kw)
        assert len(r) == 5
        for x, y in zip(r, o):
            np.testing.assert_array_equal(x, y)
    np.testing.assert_array_equal(vb.snrmap(a, 4.0, verbose=False), O.snrmap(a, 4.0))
    cube, angs = adi_cube(20, 40, 3, 60.0, seed=12, planet_peak=300.0)
    cubeout, fr, table = vb.pca(cube, angs, ncomp=(1, 4), source_xy=(28, 24), fwhm=4, verbose=False, full_output=True)
    o_cube, o_fr, o_tab, o_npc = O.pca_grid_snr(cube, angs, (1, 4), (28, 24), 4)
    assert cubeout.shape == o_cube.shape == (4, 40, 40) and list(table.columns) == ["PCs", "S/Ns", "fluxes"]
    np.testing.assert_allclose(np.asarray(table["S/Ns"]), o_tab["S/Ns"], rtol=5e-3)
    assert int(table["PCs"][int(np.argmax(table["S/Ns"]))]) == o_npc
    assert rel_err(fr, o_fr) < 3e-4
    np.testing.assert_array_equal(vb.pca(cube, angs, ncomp=(1, 4), source_xy=(28, 24), fwhm=4, verbose=False), fr)


def test_pca_annular_adimsdi_vs_oracle(vb):
    """``pca_annular(cube4d, scale_list=, ncomp=(k_ifs, k_adi))`` (``pca_local.py:332-462``, ``_pca_sdi_fr``
    :470-591): host orchestration (channel libraries per annulus, grouped Gramians, both passes, single-pass variant)
    against the oracle, which is bit-identical to the unmodified reference on these very cases
    (``test_oracle_vs_reference.py::test_pca_annular_adimsdi_bit_identical``)."""
    from tools.make_golden import ifs_cube
    cube, angs, sl = ifs_cube(z=5, n=8, size=24, seed=7)
    for ncomp, kw in (((2, 2), dict(asize=4, delta_sep=(0.1, 0.3))), ((1, None), dict(asize=6, delta_sep=0.1)),
                      ((2, 2), dict(asize=4, delta_sep=(0.05, 0.15), n_segments=2, collapse_ifs="median",
                                    scaling="temp-mean"))):
        o = O.pca_annular_sdi(cube, angs, sl, ncomp, fwhm=3, full_output=True, **kw)
        r = vb.pca_annular(cube, angs, scale_list=sl, ncomp=ncomp, fwhm=3, verbose=False, full_output=True, **kw)
        assert len(r) == 3 and r[0].shape == o[0].shape
        # the cube is float64 here and the product computes in fp32: the floor is the fp32 rounding of halo-level
        # samples through the two rescalings, eps32 * max|cube| = 5.6e-4 (measured 2.8e-4 ... 1.1e-3); the same
        # rule as the full-frame ADI+mSDI tests (a multiple of max|cube|), ten times tighter
        tol = 5e-7 * float(np.max(np.abs(cube)))
        assert np.max(np.abs(r[0] - o[0])) < tol, (ncomp, kw)
        m = ~np.isnan(o[2])
        assert np.array_equal(np.isnan(r[2]), ~m) and np.max(np.abs(r[2][m] - o[2][m])) < tol, (ncomp, kw)
    fr = vb.pca_annular(cube, angs, scale_list=sl, ncomp=(1, None), fwhm=3, asize=6, delta_sep=0.1, verbose=False)
    assert fr.shape == cube.shape[2:]
    cref, _, _ = ifs_cube(z=5, n=6, size=24, seed=8)        # reference cube through the same spectral pass
    o = O.pca_annular_sdi(cube, angs, sl, (2, 2), fwhm=3, asize=4, delta_sep=(0.1, 0.3), cube_ref=cref,
                          full_output=True)
    r = vb.pca_annular(cube, angs, scale_list=sl, ncomp=(2, 2), fwhm=3, asize=4, delta_sep=(0.1, 0.3), cube_ref=cref,
                       verbose=False, full_output=True)
    tol = 5e-7 * float(np.max(np.abs(cube)))
    assert np.max(np.abs(r[0] - o[0])) < tol
    with pytest.raises(TypeError):
        vb.pca_annular(cube, angs, scale_list=sl, ncomp=(2, 2), fwhm=3, cube_ref=cref[0], verbose=False)
    with pytest.raises(TypeError):
        vb.pca_annular(cube, angs, scale_list=sl, ncomp=2, fwhm=3, verbose=False)
    with pytest.raises(ValueError):
        vb.pca_annular(cube, angs, scale_list=sl[:-1], ncomp=(1, 1), fwhm=3, verbose=False)
    with pytest.raises(RuntimeError):                      # no channel moved by 50 FWHM
        vb.pca_annular(cube, angs, scale_list=sl, ncomp=(1, 1), fwhm=3, delta_sep=50.0, verbose=False)


def test_pca_annular_ncomp_auto_vs_oracle(vb):
    """``pca_annular(ncomp='auto')``: the noise-decay rule evaluated from Gramian eigenpairs + library row sums (what
    the kernel sees) picks the same number of components per patch as the reference's rule on the residual matrix
    (oracle pinned bit-identically: ``test_pca_annular_ncomp_auto_bit_identical``)."""
    cube, angs = adi_cube(24, 40, 4, 80.0, seed=5)
    for kw in (dict(ncomp="auto", tol=0.1, asize=5, delta_rot=(0.1, 0.4)),
               dict(ncomp="auto", tol=0.5, asize=6, n_segments=2, delta_rot=0.3),
               dict(ncomp="auto", tol=0.02, asize=5, delta_rot=0.2),
               dict(ncomp=("auto", 2, "auto", 1), tol=0.1, asize=5, delta_rot=0.2),
               dict(ncomp="auto", tol=0.1, asize=5, delta_rot=0)):
        # the rule picks 6 ... 24 components here: with that many the reference's fp32 arithmetic is itself
        # 1.2e-4 ... 3e-4 away from its own float64 evaluation, so the 1e-4 bound is taken against the float64 run
        # (measured 1e-5 ... 5e-5) and the fp32 run bounds the distance the reference's own rounding allows
        used32, used64 = [], []
        o = O.pca_annular(cube, angs, full_output=True, ncomp_out=used32, **kw)
        o64 = O.pca_annular(cube.astype(np.float64), angs, full_output=True, ncomp_out=used64, **kw)
        assert used32 == used64                     # the choice itself is not borderline on these cases
        r = vb.pca_annular(cube, angs, full_output=True, verbose=False, **kw)
        scale = np.max(np.abs(o64[0]))
        assert np.max(np.abs(r[0] - o64[0])) < 1e-4 * scale, kw
        assert np.max(np.abs(r[0] - o[0])) < 5e-4 * scale, kw
        assert rel_err(r[2], o64[2]) < 3e-4, kw


def test_detection_host_logic_vs_oracle(vb):
    """``vip_b200.detection`` (``metrics/detection.py:26-382``): mask, background level, peak ordering / spacing,
    Gaussian-fit constraints, S/N filter and return layouts against the oracle, with the kernels replaced by their
    stand-ins; the reference's ``check_detection`` criterion on the product's own PCA frame."""
    import pandas as pn
    cube, gen_angs = adi_cube(40, 64, 4, 120.0, seed=12, planet_peak=60.0)
    frame = np.nan_to_num(vb.pca(cube, -gen_angs, ncomp=4, verbose=False))
    for mode in ("lpeaks", "snrmap"):
        want = O.detection(frame, fwhm=4, mode=mode, snr_thresh=5, full_output=True)
        tab = vb.detection(frame, fwhm=4, mode=mode, snr_thresh=5, full_output=True, plot=False, verbose=False)
        assert isinstance(tab, pn.DataFrame) and list(tab.columns) == ["y", "x", "px_snr"]
        np.testing.assert_allclose(tab.y, want["y"], atol=1e-6)
        np.testing.assert_allclose(tab.x, want["x"], atol=1e-6)
        np.testing.assert_allclose(tab.px_snr, want["px_snr"], rtol=1e-6)
        assert any(abs(y - 32) <= 3 and abs(x - 51) <= 3 for y, x in zip(tab.y, tab.x))
        yy, xx = vb.detection(frame, fwhm=4, mode=mode, snr_thresh=5, plot=False, verbose=False)
        np.testing.assert_allclose(yy, want["y"], atol=1e-6)
    assert vb.detection(np.zeros((40, 40), dtype=np.float32), fwhm=4, plot=False, verbose=False) == (0, 0)
    with pytest.raises(TypeError):
        vb.detection(cube, fwhm=4, plot=False)
    with pytest.raises(ValueError):
        vb.detection(frame, fwhm=4, matched_filter=True, plot=False)
    with pytest.raises(ValueError):
        vb.detection(frame, fwhm=4, mode="blobs", plot=False)
    with pytest.raises(NotImplementedError):
        vb.detection(frame, fwhm=4, mode="log", plot=False)


def test_pca_annular_left_eigv_vs_oracle(vb):
    """``pca_annular(left_eigv=True)``: Gramian of the outside pixels as (full-frame Gramian - segment Gramian) for the
    per-pixel scalings, the outside matrix itself for 'spat-*'; against the oracle (bit-identical to the reference:
    ``test_pca_annular_left_eigv_bit_identical``)."""
    cube, angs = adi_cube(16, 32, 3, 70.0, seed=8)
    for kw in (dict(ncomp=2, asize=5), dict(ncomp=3, asize=4, n_segments=2, scaling="temp-mean"),
               dict(ncomp=2, asize=5, scaling="spat-mean"), dict(ncomp=[1, 3], asize=5)):
        if isinstance(kw["ncomp"], list):
            r = vb.pca_annular(cube, angs, left_eigv=True, verbose=False, full_output=True, **kw)
            for i, k in enumerate(kw["ncomp"]):
                o = O.pca_annular(cube, angs, left_eigv=True, full_output=True, **dict(kw, ncomp=k))
                assert np.max(np.abs(r[0][i] - o[0])) < 1e-4 * np.max(np.abs(o[0])), kw
            continue
        o = O.pca_annular(cube, angs, left_eigv=True, full_output=True, **kw)
        r = vb.pca_annular(cube, angs, left_eigv=True, verbose=False, full_output=True, **kw)
        assert np.max(np.abs(r[0] - o[0])) < 1e-4 * np.max(np.abs(o[0])), kw
        assert rel_err(r[2], o[2]) < 3e-4, kw
    with pytest.raises(NotImplementedError):
        vb.pca_annular(cube, angs, left_eigv=True, ncomp="auto", verbose=False)


def _ifs_small():
    from tools.make_golden import ifs_cube
    cube, angs, sl = ifs_cube(z=5, n=10, size=24, seed=43)
    cref = ifs_cube(z=5, n=6, size=24, seed=47)[0]
    return cube, angs, sl, cref


def test_pca_adimsdi_fullframe_options(vb):
    """Full-frame ADI+mSDI options added in the second half of round 2, against the oracle restatements that are
    pinned bit-identically to the unmodified reference (tests/test_oracle_vs_reference.py): reference cube in both
    modes, PA-rejection second pass (``source_xy``) with ``cube_sig``, single-pass PCA grid with / without
    ``source_xy``.  Tolerance as for the other mSDI tests: 5e-6 * max|cube| (fp32 pipeline vs float64 reference)."""
    cube, angs, sl, cref = _ifs_small()
    tol = 5e-6 * float(np.max(np.abs(cube)))
    kw = dict(scale_list=sl, verbose=False)
    err = lambda a, b: float(np.max(np.abs(np.asarray(a, dtype=np.float64) - b)))      # noqa: E731
    # double pass with a reference cube (RSDI): the first-pass frames of the reference come back too
    fr, rc, rd = vb.pca(cube, angs, cube_ref=cref, adimsdi="double", ncomp=(2, 3), full_output=True, **kw)
    ofr, orc, ord_ = O.pca_adimsdi_double(cube, angs, sl, (2, 3), cube_ref=cref, full_output=True)
    assert fr.dtype == np.float64 and rc.shape == orc.shape == (16, 24, 24) and rd.shape == ord_.shape
    assert err(rc, orc) < tol and err(fr, ofr) < tol
    with pytest.raises(IndexError):
        vb.pca(cube, angs, cube_ref=cref, adimsdi="double", ncomp=(2, 3), ref_strategy="ARDI", **kw)
    # double pass, PA-rejection libraries in the second pass
    sig = np.zeros(cube.shape[1:], dtype=np.float32)
    sig[:, 12, 17] = 3.0
    for extra in (dict(), dict(cube_ref=cref), dict(cube_sig=sig)):
        opts = dict(source_xy=(17, 12), delta_rot=0.5, fwhm=3, min_frames_pca=2)
        fr, rc, rd = vb.pca(cube, angs, adimsdi="double", ncomp=(2, 2), full_output=True, **opts, **kw, **extra)
        ofr, orc, ord_ = O.pca_adimsdi_double(cube, angs, sl, (2, 2), full_output=True, **opts, **extra)
        m = ~np.isnan(ord_)
        assert err(rc, orc) < tol and err(rd[m], ord_[m]) < 4 * tol and err(fr, ofr) < 4 * tol, list(extra)
    # single pass with a reference cube, both strategies
    for strat, tag in (("RDI", "RSDI"), ("ARDI", "ARSDI")):
        fr, allfr, desc, resadi = vb.pca(cube, angs, cube_ref=cref, adimsdi="single", ncomp=3, ref_strategy=strat,
                                         full_output=True, **kw)
        ofr, oall, odesc, oadi = O.pca_adimsdi_single(cube, angs, sl, 3, cube_ref=cref, ref_strategy=tag,
                                                      full_output=True)
        assert allfr.shape == oall.shape and desc.shape == odesc.shape and desc.dtype == np.float32
        assert err(allfr, oall) < tol and err(resadi, oadi) < tol and err(fr, ofr) < tol, strat
    # single-pass grid
    out, pcl = vb.pca(cube, angs, adimsdi="single", ncomp=(1, 3), full_output=True, **kw)
    oout, opcl = O.pca_adimsdi_single_grid(cube, angs, sl, (1, 3))
    assert pcl == opcl and out.dtype == np.float64 and out.shape == oout.shape and err(out, oout) < tol
    out = vb.pca(cube, angs, adimsdi="single", ncomp=[2, 4], ifs_collapse_range=(1, 4), collapse="mean", **kw)
    assert err(out, O.pca_adimsdi_single_grid(cube, angs, sl, [2, 4], ifs_collapse_range=(1, 4), collapse="mean")[0]) < tol
    out = vb.pca(cube, angs, adimsdi="single", ncomp=(1, 3), med_of_npcs=True, **kw)
    assert out.shape == (24, 24) and err(out, np.median(oout, axis=0)) < tol
    out, best, table = vb.pca(cube, angs, adimsdi="single", ncomp=(1, 3), source_xy=(17, 12), fwhm=3,
                              full_output=True, **kw)
    o = O.pca_adimsdi_single_grid(cube, angs, sl, (1, 3), source_xy=(17, 12), fwhm=3)
    assert err(out, o[0]) < tol and list(table["PCs"]) == o[2]["PCs"]
    assert np.allclose(table["S/Ns"], o[2]["S/Ns"], rtol=2e-3, atol=2e-4)
    assert np.allclose(table["fluxes"], o[2]["fluxes"], rtol=1e-4, atol=10 * tol)
    assert err(best, o[0][int(np.argmax(table["S/Ns"]))]) < tol
    best2 = vb.pca(cube, angs, adimsdi="single", ncomp=(1, 3), source_xy=(17, 12), fwhm=3, **kw)
    assert np.array_equal(best, best2)


def _rdi_masks(size):
    yy, xx = O.get_annulus_segments((size, size), 3, 7)[0]
    boat = np.zeros((size, size)); boat[yy, xx] = 1
    yy, xx = O.get_annulus_segments((size, size), 8, 6)[0]
    anchor = np.zeros((size, size)); anchor[yy, xx] = 1
    return anchor, boat


def test_pca_mask_rdi_data_imputation(vb):
    """``pca(cube, angs, cube_ref=, mask_rdi=(anchor, boat) | anchor, ncomp=k)`` against the oracle restatement of
    ``cube_subtract_sky_pca`` that is pinned bit-identically to the unmodified reference
    (tests/test_oracle_vs_reference.py::test_pca_mask_rdi_bit_identical); the reference's fp32 arithmetic (sgemm
    Gramian of a 1e4 halo, fp32 matrix inverse) is the noisy side, so the float64 run of the oracle is the truth."""
    cube, angs = adi_cube(16, 33, 3, 60.0, seed=21)
    cref = adi_cube(12, 33, 3, 60.0, seed=22)[0]
    anchor, boat = _rdi_masks(33)
    for masks in ((anchor, boat), anchor):
        o64 = O.pca_fullframe(cube.astype(np.float64), angs, cube_ref=cref.astype(np.float64), mask_rdi=masks, ncomp=3,
                              full_output=True)
        fr, pcs, recon, res, res_ = vb.pca(cube, angs, cube_ref=cref, mask_rdi=masks, ncomp=3, verbose=False,
                                           full_output=True)
        assert fr.dtype == np.float32 and pcs.shape == o64[1].shape == (12, 33, 33) and recon.shape == cube.shape
        assert np.max(np.abs(res - o64[3])) < 1e-4 * np.max(np.abs(o64[3]))
        assert np.max(np.abs(recon - o64[2])) < 1e-5 * np.max(np.abs(o64[2]))
        for j in range(3):                                   # boat components up to their sign
            d = min(np.max(np.abs(pcs[j] - o64[1][j])), np.max(np.abs(pcs[j] + o64[1][j])))
            assert d < 1e-4 * np.max(np.abs(o64[1][j])), j
        assert rel_err(fr, o64[0]) < 3e-4
        assert np.array_equal(vb.pca(cube, angs, cube_ref=cref, mask_rdi=masks, ncomp=3, verbose=False), fr)
    with pytest.raises(TypeError):
        vb.pca(cube, angs, cube_ref=cref, mask_rdi=(anchor, boat), ncomp=3, ref_strategy="ARDI", verbose=False)
    with pytest.raises(TypeError):
        vb.pca(cube, angs, mask_rdi=(anchor, boat), ncomp=3, verbose=False)


def test_pca_batch_from_fits_paths(vb, tmp_path):
    """``pca(cube='cube.fits', angle_list='angs.fits', batch=...)`` (``utils_pca.py:508-519``): the mini-batches are
    sliced from a memory map of the file's big-endian data unit; same result as with the arrays."""
    cube, angs = adi_cube(23, 33, 3, 60.0, seed=11)
    cpath, apath = str(tmp_path / "cube.fits"), str(tmp_path / "angs.fits")
    vb.write_fits(cpath, cube, verbose=False)
    vb.write_fits(apath, angs.astype(np.float32), verbose=False)
    want = vb.pca(cube, angs.astype(np.float32).astype(np.float64), ncomp=3, batch=6, verbose=False, full_output=True)
    got = vb.pca(cpath, apath, ncomp=3, batch=6, verbose=False, full_output=True)
    for a, b in zip(got, want):
        np.testing.assert_array_equal(a, b)


def test_pca_annular_beyond_the_batched_solver_limits(vb):
    """``pca_annular`` with more components (ncomp > 24) or larger libraries (> 256 frames) than the batched
    sub-Gramian eigensolver takes: the frame-by-frame route through the full-size eigensolvers, against the oracle
    on the float64-cast cube (the reference's fp32 arithmetic with 30 components of a 36-frame library is the noisy
    side)."""
    cube, angs = adi_cube(40, 36, 3, 80.0, seed=9)
    kw = dict(ncomp=30, asize=6, delta_rot=0.05, radius_int=2)
    co, cd, fr = vb.pca_annular(cube, angs, verbose=False, full_output=True, **kw)
    oo, od, of = O.pca_annular(cube.astype(np.float64), angs, full_output=True, **kw)
    assert np.max(np.abs(co - oo)) < 1e-4 * np.max(np.abs(oo))
    assert rel_err(fr, of) < 3e-4
    # libraries of 270 frames
    cube, angs = adi_cube(300, 20, 3, 170.0, seed=10)
    kw = dict(ncomp=4, asize=5, delta_rot=0.1, max_frames_lib=270, radius_int=2)
    co, cd, fr = vb.pca_annular(cube, angs, verbose=False, full_output=True, **kw)
    oo, od, of = O.pca_annular(cube.astype(np.float64), angs, full_output=True, **kw)
    assert np.max(np.abs(co - oo)) < 1e-4 * np.max(np.abs(oo))
    assert rel_err(fr, of) < 3e-4
