"""Generate tests/golden/*.npz by running the UNMODIFIED reference (vip_hci from /root/reference).

Run in the build container only:  python tools/make_golden.py
Inputs are regenerated from seeds by tools/synth.py (so only outputs are stored, as fp32/fp64 as the
reference returns them).  The fixtures pin oracle/vip_oracle.py and the CUDA path.
"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.simplefilter("ignore")

from oracle import ref_loader          # noqa: E402
from tools.synth import adi_cube        # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def golden_inputs():
    """Seeded inputs shared by make_golden.py and the tests."""
    rng = np.random.default_rng(777)
    d = {}
    # derotation: even/odd sizes, angles covering all rot90 quadrants, the 135/315 rint quirk, NaNs
    angs = np.array([12.3, -33.0, 77.7, 181.0, 300.5, 135.0, 315.0, 45.0, -400.2, 90.0])
    for S in (32, 33):
        cube = rng.normal(size=(10, S, S)).astype(np.float32)
        cube[2, 3, 4] = np.nan
        cube[5, S - 1, 0] = np.nan
        d[f"derot{S}"] = (cube, angs)
    cube, a = adi_cube(12, 128, 4, 70.0, seed=31)          # FFT path (N=512)
    d["derot128"] = (cube - cube.mean(0, keepdims=True), a * 3.7)
    d["c1"] = adi_cube(50, 101, 5, 60.0, seed=20260102)    # BASELINE config 1
    d["small"] = adi_cube(30, 41, 4, 60.0, seed=5)
    d["ann"] = adi_cube(40, 48, 3, 80.0, seed=9)
    d["ifs"] = ifs_cube()
    return d


def ifs_cube(z=6, n=14, size=32, seed=41):
    """Small IFS cube (z, n, size, size): each channel = the same ADI scene radially stretched by
    lambda/lambda_min (speckles scale with wavelength), plus noise.  Returns (cube, angles, scale_list)."""
    rng = np.random.default_rng(seed)
    lam = np.linspace(1.0, 1.4, z)
    scale_list = lam.max() / lam
    base, angs = adi_cube(n, 2 * size, 3, 50.0, seed=seed)
    from oracle import vip_oracle as O
    cube = np.empty((z, n, size, size), dtype=np.float32)
    c0 = size // 2
    for c in range(z):
        for i in range(n):
            fr = O.frame_rescaling(base[i].astype(np.float64), lam[c] / lam[0])
            cube[c, i] = fr[size - c0:size - c0 + size, size - c0:size - c0 + size]
    cube += rng.normal(scale=0.5, size=cube.shape).astype(np.float32)
    return cube, angs, scale_list


def main():
    vip = ref_loader.load()
    from vip_hci.psfsub import pca, pca_annular
    from vip_hci.preproc import cube_derotate, cube_collapse
    os.makedirs(OUT, exist_ok=True)
    inp = golden_inputs()
    out = {}
    for key in ("derot32", "derot33", "derot128"):
        cube, angs = inp[key]
        out[key] = cube_derotate(cube, angs)
    cube, angs = inp["derot32"]
    out["derot32_mask0"] = cube_derotate(np.where(np.isnan(cube), 0, cube), angs, mask_val=0, interp_zeros=True)
    np.savez_compressed(os.path.join(OUT, "derotate.npz"), **out)

    out = {}
    cube, angs = inp["c1"]
    fr, pcs, recon, res, res_ = pca(cube, angs, ncomp=5, verbose=False, full_output=True)
    out["c1_frame"] = fr
    out["c1_res_frame7"] = res[7]
    out["c1_resder_frame7"] = res_[7]
    out["c1_proj"] = (pcs.reshape(5, -1).T @ pcs.reshape(5, -1)).astype(np.float32)[::97, ::89]  # projector samples
    cube, angs = inp["small"]
    for mode in ("lapack", "eigen"):
        out[f"small_{mode}"] = pca(cube, angs, ncomp=4, svd_mode=mode, verbose=False)
    for sc in ("temp-mean", "spat-mean", "temp-standard", "spat-standard"):
        out[f"small_{sc}"] = pca(cube, angs, ncomp=3, scaling=sc, verbose=False)
    for col in ("mean", "sum"):
        out[f"small_{col}"] = pca(cube, angs, ncomp=3, collapse=col, verbose=False)
    ref = adi_cube(20, 41, 4, 60.0, seed=6)[0]
    out["small_rdi"] = pca(cube, angs, cube_ref=ref, ncomp=4, verbose=False)
    out["small_ardi"] = pca(cube, angs, cube_ref=ref, ncomp=4, ref_strategy="ARDI", verbose=False)
    out["small_cevr"] = pca(cube, angs, ncomp=0.9995, verbose=False)
    np.savez_compressed(os.path.join(OUT, "pca_fullframe.npz"), **out)
    make_left_eigv(inp, pca)

    out = {}
    cube, angs = inp["ann"]
    co, cd, fr = pca_annular(cube, angs, ncomp=3, asize=6, verbose=False, full_output=True)
    out["ann_frame"], out["ann_cube_out5"] = fr, co[5]
    out["ann_seg_frame"] = pca_annular(cube, angs, ncomp=2, asize=6, n_segments=3, delta_rot=0.5,
                                       radius_int=4, verbose=False)
    co, cd, frl = pca_annular(cube, angs, ncomp=[1, 3], asize=6, verbose=False, full_output=True)
    out["ann_list_frames"], out["ann_list_cube_out_1_5"] = np.array(frl), co[1, 5]
    np.savez_compressed(os.path.join(OUT, "pca_annular.npz"), **out)

    out = {}
    cube, angs, sl = inp["ifs"]
    fr, rc, rd = pca(cube, angs, scale_list=sl, adimsdi="double", ncomp=(2, 3), verbose=False, full_output=True)
    out["double_frame"], out["double_res_channels"], out["double_res_der"] = fr, rc, rd
    out["double_skipadi"] = pca(cube, angs, scale_list=sl, adimsdi="double", ncomp=(2, None), verbose=False)
    out["double_range"] = pca(cube, angs, scale_list=sl, adimsdi="double", ncomp=(1, 2), ifs_collapse_range=(1, 5),
                              collapse_ifs="median", verbose=False)
    np.savez_compressed(os.path.join(OUT, "pca_sdi.npz"), **out)

    out = {}
    cube = inp["small"][0].copy()
    cube[3, 5, 5] = np.nan
    cube[:, 7, 7] = np.nan
    for m in ("median", "mean", "sum", "max", "absmean", "trimmean"):
        out[m] = cube_collapse(cube.copy(), m, n=10)
    out["median_even"] = cube_collapse(cube[:-1].copy(), "median")
    w = np.random.default_rng(1).uniform(size=30)
    out["wmean"] = cube_collapse(cube.copy(), "wmean", w=w)
    np.savez_compressed(os.path.join(OUT, "collapse.npz"), **out)
    make_shift_medsub(inp)
    make_source_xy(inp, pca)
    make_grid_4d(inp, pca)
    make_incremental(inp, pca)
    make_sdi_single(inp, pca)
    make_annular_4d(inp, pca_annular)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


def shift_inputs():
    """Seeded shift vectors / cubes shared by make_golden.py and the tests."""
    rng = np.random.default_rng(4242)
    d = {}
    d["odd"] = (rng.normal(size=(7, 41, 41)).astype(np.float32), rng.uniform(-4, 4, 7), rng.uniform(-4, 4, 7))
    sy = np.array([0.0, 1.0, -2.5, 0.3, 3.999, -0.001])
    sx = np.array([0.0, -3.0, 1.5, -0.7, 0.5, 2.2])
    d["even"] = (rng.normal(size=(6, 32, 32)).astype(np.float32), sy, sx)
    d["rect"] = (rng.normal(size=(4, 20, 24)), rng.uniform(-2, 2, 4), rng.uniform(-2, 2, 4))     # float64
    return d


def make_shift_medsub(inp):
    """cube_shift (vip-fft) and full-frame median_sub."""
    ref_loader.load()
    from vip_hci.preproc import cube_shift
    from vip_hci.psfsub import median_sub
    out = {}
    for key, (cube, sy, sx) in shift_inputs().items():
        out[f"shift_{key}"] = cube_shift(cube, sy, sx, nproc=1)
    out["shift_scalar"] = cube_shift(shift_inputs()["even"][0], 1.25, -0.75, nproc=1)
    cube, angs = inp["small"]
    co, cd, fr = median_sub(cube, angs, verbose=False, full_output=True)
    out["med_cube_out3"], out["med_cube_der3"], out["med_frame"] = co[3], cd[3], fr
    out["med_mean"] = median_sub(cube, angs, collapse="mean", verbose=False)
    ref = adi_cube(20, 41, 4, 60.0, seed=6)[0]
    out["med_rdi_median"] = median_sub(cube, angs, cube_ref=ref, verbose=False)
    out["med_rdi_mean"] = median_sub(cube, angs, cube_ref=ref, collapse_ref="mean", verbose=False)
    np.savez_compressed(os.path.join(OUT, "shift_medsub.npz"), **out)


def make_left_eigv(inp, pca):
    """pca(..., left_eigv=True): projection on the temporal singular vectors."""
    cube, angs = inp["small"]
    fr, pcs, recon, res, res_ = pca(cube, angs, ncomp=4, left_eigv=True, verbose=False, full_output=True)
    out = {"left_frame": fr, "left_pcs": pcs, "left_res": res}
    out["left_scaled_frame"] = pca(cube, angs, ncomp=3, left_eigv=True, scaling="spat-mean", verbose=False)
    np.savez_compressed(os.path.join(OUT, "pca_left_eigv.npz"), **out)


SOURCE_XY_CASES = {
    "plain": dict(source_xy=(33, 28), delta_rot=0.5, fwhm=4, ncomp=3, min_frames_pca=5),
    "trunc": dict(source_xy=(33, 28), delta_rot=0.3, fwhm=4, ncomp=3, max_frames_pca=12, min_frames_pca=5),
    "rdi": dict(source_xy=(28, 12), delta_rot=1, fwhm=4, ncomp=4),          # + cube_ref
    "scaled": dict(source_xy=(33, 28), delta_rot=0.5, fwhm=4, ncomp=2, scaling="temp-mean", min_frames_pca=5),
}


def make_source_xy(inp, pca):
    """pca(..., source_xy=..., delta_rot=...): frame-by-frame PCA with a PA-rejection library."""
    cube, angs = inp["small"]
    ref = adi_cube(20, 41, 4, 60.0, seed=6)[0]
    out = {}
    for key, kw in SOURCE_XY_CASES.items():
        extra = dict(cube_ref=ref) if key == "rdi" else {}
        fr, recon, res, res_ = pca(cube, angs, verbose=False, full_output=True, **kw, **extra)
        out[f"{key}_frame"], out[f"{key}_res"], out[f"{key}_recon7"] = fr, res, recon[7]
    np.savez_compressed(os.path.join(OUT, "pca_source_xy.npz"), **out)


INCREMENTAL_CASES = {"b12": dict(batch=12, ncomp=4), "b17_mean": dict(batch=17, ncomp=3, collapse="mean")}


def make_incremental(inp, pca):
    """pca(..., batch=...): incremental PCA in mini-batches (scikit-learn IncrementalPCA), median of batch frames."""
    cube, angs = inp["small"]
    out = {}
    for key, kw in INCREMENTAL_CASES.items():
        fr, pcs, med = pca(cube, angs, verbose=False, full_output=True, **kw)
        out[f"{key}_frame"], out[f"{key}_pcs"], out[f"{key}_medians"] = fr, pcs, med
    np.savez_compressed(os.path.join(OUT, "pca_incremental.npz"), **out)


def make_annular_4d(inp, pca_annular):
    """pca_annular on a 4-d cube without scale_list (per-channel annular ADI + collapse_ifs)."""
    cube4, angs4, _ = inp["ifs"]
    cube4 = cube4[:3]
    out = {}
    co, cd, fr = pca_annular(cube4, angs4, ncomp=2, asize=5, delta_rot=(0.05, 0.2), verbose=False, full_output=True)
    out["ann4d_frame"], out["ann4d_cube_out_ch1_fr3"], out["ann4d_cube_der_ch2_fr5"] = fr, co[1, 3], cd[2, 5]
    out["ann4d_list_median"] = pca_annular(cube4, angs4, ncomp=[1, 2, 3], asize=5, delta_rot=0.1,
                                           collapse_ifs="median", verbose=False)
    np.savez_compressed(os.path.join(OUT, "pca_annular_4d.npz"), **out)


def make_sdi_single(inp, pca):
    """ADI+mSDI single-pass PCA (one PCA over all rescaled channels of all frames)."""
    cube, angs, sl = inp["ifs"]
    out = {}
    fr, allfr, desc, resadi = pca(cube, angs, scale_list=sl, adimsdi="single", ncomp=3, verbose=False,
                                  full_output=True)
    out["single_frame"], out["single_allfr_7"], out["single_desc_ch2"], out["single_resadi"] = \
        fr, allfr[7], desc[2], resadi
    out["single_nocrop"] = pca(cube, angs, scale_list=sl, adimsdi="single", ncomp=2, crop_ifs=False,
                               collapse_ifs="median", verbose=False)
    out["single_range"] = pca(cube, angs, scale_list=sl, adimsdi="single", ncomp=2, ifs_collapse_range=(1, 5),
                              verbose=False)
    np.savez_compressed(os.path.join(OUT, "pca_sdi_single.npz"), **out)


def make_grid_4d(inp, pca):
    """PCA grid (tuple/list ncomp -> pca_grid) and 4-d cubes without scale_list (per-channel ADI)."""
    out = {}
    cube, angs = inp["small"]
    fr, pcl = pca(cube, angs, ncomp=(1, 4), verbose=False, full_output=True)
    out["grid_range"], out["grid_range_pclist"] = fr, np.asarray(pcl)
    out["grid_list"] = pca(cube, angs, ncomp=[2, 4], verbose=False)
    out["grid_step_med"] = pca(cube, angs, ncomp=(1, 5, 2), med_of_npcs=True, verbose=False)
    ref = adi_cube(20, 41, 4, 60.0, seed=6)[0]
    out["grid_rdi"] = pca(cube, angs, cube_ref=ref, ncomp=(2, 3), scaling="temp-mean", verbose=False)
    cube4, angs4, _ = inp["ifs"]
    r = pca(cube4, angs4, ncomp=2, verbose=False, full_output=True)
    out["ch_frame"], out["ch_pcs"], out["ch_res_der"], out["ch_ifs"] = r[0], r[1], r[4], r[5]
    out["ch_list"] = pca(cube4, angs4, ncomp=[2, 2, 2, 2, 2, 2], collapse_ifs="median", verbose=False)
    g = pca(cube4, angs4, ncomp=[1, 3], verbose=False, full_output=True)
    out["ch_grid"], out["ch_grid_ifs"] = g[0], g[2]
    np.savez_compressed(os.path.join(OUT, "pca_grid4d.npz"), **out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "shift_medsub":      # regenerate one fixture file only
        os.makedirs(OUT, exist_ok=True)
        make_shift_medsub(golden_inputs())
    elif len(sys.argv) > 1 and sys.argv[1] == "left_eigv":
        ref_loader.load()
        from vip_hci.psfsub import pca as _pca
        make_left_eigv(golden_inputs(), _pca)
    elif len(sys.argv) > 1 and sys.argv[1] == "incremental":
        ref_loader.load()
        from vip_hci.psfsub import pca as _pca
        make_incremental(golden_inputs(), _pca)
    elif len(sys.argv) > 1 and sys.argv[1] == "source_xy":
        ref_loader.load()
        from vip_hci.psfsub import pca as _pca
        make_source_xy(golden_inputs(), _pca)
    else:
        main()
