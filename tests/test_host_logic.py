"""Host-side behaviour of the drop-in boundary that needs no GPU: argument parsing, validation
errors (same exception types as the reference), no silent CPU fallback."""
import numpy as np
import pytest

import vip_b200
from vip_b200.config import check_array, separate_kwargs_dict, setup_parameters, SvdMode
from vip_b200.psfsub.pca_fullfr import PCA_Params
from vip_b200.psfsub.annular import PCA_ANNULAR_Params


def test_params_field_order_matches_reference():
    names = list(PCA_Params.__dataclass_fields__)
    assert names[:8] == ["cube", "angle_list", "cube_ref", "scale_list", "ncomp", "svd_mode", "scaling",
                         "mask_center_px"]
    assert names[-6:] == ["weights", "left_eigv", "min_frames_pca", "max_frames_pca", "cube_sig", "med_of_npcs"]
    assert len(names) == 34
    ann = list(PCA_ANNULAR_Params.__dataclass_fields__)
    assert ann[:11] == ["cube", "angle_list", "cube_ref", "scale_list", "radius_int", "fwhm", "asize",
                        "n_segments", "delta_rot", "delta_sep", "ncomp"]
    assert len(ann) == 28
    p = PCA_Params(None, None, None, None, 7, "eigen")
    assert p.ncomp == 7 and p.svd_mode == SvdMode.EIGEN == "eigen"


def test_separate_kwargs():
    mine, rest = separate_kwargs_dict({"ncomp": 3, "mask_val": 0, "algo_params": 1, "imlib": "vip-fft"}, PCA_Params)
    assert mine == {"ncomp": 3, "imlib": "vip-fft"} and rest == {"mask_val": 0, "algo_params": 1}

    def f(cube, ncomp, full_output, other=1):
        pass
    got = setup_parameters(PCA_Params(cube=1, ncomp=2, full_output=False), f, full_output=True)
    assert got == {"cube": 1, "ncomp": 2, "full_output": True}


def test_check_array():
    check_array(np.zeros((2, 3, 4)), (3, 4))
    check_array([1, 2], 1)
    with pytest.raises(TypeError):
        check_array(np.zeros((3, 4)), (3, 4), msg="cube")
    with pytest.raises(TypeError):
        check_array("x", 3)
    with pytest.raises(ValueError):
        check_array(np.zeros(3), 7)


def test_pca_validation_errors_before_any_gpu_work():
    cube = np.zeros((5, 8, 8), np.float32)
    angs = np.arange(5.0)
    with pytest.raises(TypeError):
        vip_b200.pca(np.zeros((8, 8)), angs)
    with pytest.raises(NotImplementedError):
        vip_b200.pca(cube, angs, left_eigv=True, cube_ref=cube)
    with pytest.raises(NotImplementedError):
        vip_b200.pca(cube, angs, imlib="opencv")
    with pytest.raises(NotImplementedError):
        vip_b200.pca(np.zeros((2, 5, 8, 8), np.float32), angs, batch=2)      # incremental PCA: 3-d cubes only
    with pytest.raises(ValueError):
        vip_b200.pca(cube, angs, batch=2, cube_ref=cube)                     # "RDI not compatible with batch mode"
    with pytest.raises(TypeError):
        vip_b200.pca(cube, angs, cube_ref=cube, ref_strategy="XYZ")
    with pytest.raises(TypeError):
        vip_b200.cube_derotate(np.zeros((8, 8)), angs)
    with pytest.raises(NotImplementedError):
        vip_b200.cube_derotate(cube, angs, imlib="opencv")
    with pytest.raises(ValueError):
        vip_b200.cube_derotate(cube, angs, cxy=(1, 1))


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    cube = np.ones((5, 8, 8), np.float32)
    with pytest.raises(RuntimeError, match="CUDA"):
        vip_b200.pca(cube, np.arange(5.0), ncomp=1, verbose=False)
    with pytest.raises(RuntimeError, match="CUDA"):
        vip_b200.cube_derotate(cube, np.arange(5.0))


def test_product_never_imports_oracle():
    import os
    import re
    root = os.path.dirname(vip_b200.__file__)
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f


def test_rotation_scalars_vectorised_equals_scalar_form():
    """Per-frame rot90 count and shear coefficients: the vectorised host code gives the bits of the
    reference's frame-by-frame expressions (derotation.py:570-603), quirk angles included."""
    from vip_b200.preproc.derotation import rotation_scalars, _rotation_scalars_loop
    rng = np.random.default_rng(0)
    angs = np.concatenate([rng.uniform(-800, 800, 5000),
                           [0, 45, 90, 135, 180, 225, 270, 315, 360, -45, -90, 44.999999, 45.0000001, 720, -720]])
    k0, a0, b0 = _rotation_scalars_loop(angs)
    k1, a1, b1 = rotation_scalars(angs)
    np.testing.assert_array_equal(k0, k1)
    np.testing.assert_array_equal(a0, a1)
    np.testing.assert_array_equal(b0, b1)


def test_bench_reference_arm_contract():
    """``bench.py --impl reference`` (the driver's CPU arm): one JSON line with the contract's keys, runnable
    without a GPU; non-zero ranks of a torchrun launch print nothing."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
           "--config", "small"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["value"] > 0 and d["vs_baseline"] is None
    have_ref = os.path.isdir(os.path.join(root, "baseline", "_ref", "vip_hci"))
    assert d["cpu_baseline"]["kind"] == ("reference" if have_ref else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    # ms_per_step is the MEASURED time of a sample (it has to fit in the driver's wall clock around the run)
    assert d["ms_per_step"] * 1e-3 < 120 and d["extrapolation"]["full_workload_seconds"] > 0
    assert set(d["config"]) == {"workload", "l2_policy", "multi_gpu"}      # same object as the GPU arm's
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=root, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_standard_scaling_zero_variance_rule_is_sklearns():
    """'*-standard' scaling on low-flux data: sklearn treats a std below 10*eps(dtype) -- an ABSOLUTE threshold,
    1.2e-6 for float32 cubes -- as zero variance and leaves the column unscaled (var/shapes.py:740-781 ->
    sklearn.preprocessing.scale -> _handle_zeros_in_scale).  Pure torch host logic: runs on CPU tensors."""
    import torch
    import warnings
    from oracle import vip_oracle as O
    from vip_b200.psfsub.pca_fullfr import scale_matrix_device
    rng = np.random.default_rng(3)
    n = 40
    M = np.empty((n, 6), dtype=np.float64)
    for j, sd in enumerate((0.0, 1e-7, 5e-7, 3e-6, 1e-3, 2.0)):       # around the 1.19e-6 threshold
        M[:, j] = rng.standard_normal(n) * sd
    M32 = M.astype(np.float32)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for mode in ("temp-standard", "spat-standard", "temp-mean", "spat-mean"):
            ref = O.matrix_scaling(M32, mode)
            out = scale_matrix_device(torch.from_numpy(M32), mode, np.float32).numpy()
            np.testing.assert_allclose(out, ref, rtol=2e-5, atol=1e-9)
        # a float64 cube: the threshold follows the input dtype (2.2e-15), so the 1e-7 column IS scaled
        ref64 = O.matrix_scaling(M, "temp-standard")
        out64 = scale_matrix_device(torch.from_numpy(M32), "temp-standard", np.float64).numpy()
        assert abs(np.std(ref64[:, 1]) - 1) < 1e-6 and abs(np.std(out64[:, 1]) - 1) < 1e-3
        assert np.std(O.matrix_scaling(M32, "temp-standard")[:, 1]) < 1e-6
