"""The reference's OWN offline unit tests for the host-side pieces of the path, executed verbatim from
/root/reference/tests with this repo's host functions substituted for the reference's (build container only;
skipped where the checkout is absent).  Nothing is copied: the test modules are loaded from where they lie,
their known-answer arrays stay in the reference tree (last bullet / section 8c).

Substituted (all host numpy code, no GPU): ``vip_b200.var`` (frame_center, dist, mask_circle,
get_annulus_segments, reshape_matrix), ``vip_b200.preproc`` (_find_indices_adi, _define_annuli,
check_scal_vector, _find_indices_sdi) and, where the product computes on the GPU, the oracle's restatement
(matrix_scaling, cube_rescaling_wavelengths, get_square).
"""
import importlib.util
import os
import sys
import types
import warnings

import numpy as np
import pytest

from oracle import ref_loader, vip_oracle as O

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference checkout not present")

REF_TESTS = os.path.join(os.path.dirname(ref_loader.REFERENCE_SRC.rstrip("/")), "tests")


def _helpers_module():
    """Stand-in for the reference's ``tests/helpers.py`` (which imports ratelimit / astropy / requests at module
    scope): only the names the unit-test modules use."""
    m = types.ModuleType("tests.helpers")

    def aarc(actual, desired, rtol=1e-5, atol=1e-6):
        np.testing.assert_allclose(actual, desired, rtol=rtol, atol=atol)
    m.aarc, m.np = aarc, np
    m.param, m.raises, m.fixture, m.mark = pytest.param, pytest.raises, pytest.fixture, pytest.mark
    m.parametrize, m.filterwarnings = pytest.mark.parametrize, pytest.mark.filterwarnings
    m.check_detection = m.download_resource = None
    return m


def _load(relpath):
    ref_loader.load()
    saved = {k: sys.modules.get(k) for k in ("tests", "tests.helpers")}
    pkg = types.ModuleType("tests")
    pkg.__path__ = []
    pkg.helpers = _helpers_module()
    sys.modules["tests"], sys.modules["tests.helpers"] = pkg, pkg.helpers
    try:
        name = "ref_unit_" + os.path.basename(relpath)[:-3]
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF_TESTS, relpath))
        mod = importlib.util.module_from_spec(spec)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


def _cases(func):
    """(args tuple) for every parametrisation of a reference test function (plain call if none)."""
    out = [()]
    for mark in getattr(func, "pytestmark", []):
        if mark.name == "parametrize":
            names = [s.strip() for s in mark.args[0].split(",")]
            vals = []
            for v in mark.args[1]:
                v = getattr(v, "values", v)
                vals.append(tuple(v) if len(names) > 1 else (v,))
            out = [a + b for a in out for b in vals]
    return out


def _run(mod, test_name, select=None):
    func = getattr(mod, test_name)
    n = 0
    for args in _cases(func):
        if select is not None and not select(args):
            continue
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            func(*args)
        n += 1
    assert n > 0
    return n


def test_var_shapes_known_answers():
    """tests/pre_3_10/test_var_shapes.py:49-82, 196-479, 504-545."""
    import vip_b200.var.shapes as shp
    import vip_b200.var.coords as crd
    mod = _load("pre_3_10/test_var_shapes.py")
    ref_seg = mod.get_annulus_segments

    def segments(data, inner_radius, width, nsegm=1, theta_init=0, optim_scale_fact=1, mode="ind", out=False):
        # index mode (the one the path uses) is ours; 'val' / 'mask' / optim_scale_fact stay the reference's
        if mode != "ind" or optim_scale_fact != 1 or out:
            return ref_seg(data, inner_radius, width, nsegm, theta_init, optim_scale_fact, mode, out)
        return shp.get_annulus_segments(data, inner_radius, width, nsegm, theta_init)
    mod.frame_center, mod.dist = crd.frame_center, crd.dist
    mod.mask_circle, mod.reshape_matrix = shp.mask_circle, shp.reshape_matrix
    mod.get_annulus_segments = segments
    mod.matrix_scaling = O.matrix_scaling
    for t in ("test_frame_center", "test_mask_circle", "test_get_annulus_segments", "test_dist",
              "test_reshape_matrix", "test_matrix_scaling"):
        _run(mod, t)


def test_rotation_helpers_known_answers():
    """tests/pre_3_10/test_preproc_rotation.py:131-186 (``_define_annuli`` 53.13 deg, seven ``_find_indices_adi``
    index lists incl. truncation)."""
    import vip_b200.preproc.derotation as d
    mod = _load("pre_3_10/test_preproc_rotation.py")
    mod._find_indices_adi, mod._define_annuli = d._find_indices_adi, d._define_annuli
    _run(mod, "test_define_annuli")
    assert _run(mod, "test_find_indices_adi") >= 7


def test_rotation_24_steps_identity_oracle():
    """tests/pre_3_10/test_preproc_rotation.py:21-69, the vip-fft / constant-border / no-edge-blend case (the options
    the path supports), with the oracle's rotation (the product's rotation is a GPU kernel, pinned to the oracle
    in tests/test_gpu_parity.py; ``edge_blend`` and other border modes need astropy and are out of scope)."""
    mod = _load("pre_3_10/test_preproc_rotation.py")

    def derot(array, angle_list, imlib="vip-fft", interpolation=None, nproc=1, border_mode="constant",
              edge_blend=None, **kw):
        assert imlib == "vip-fft" and border_mode == "constant" and edge_blend is None
        return O.cube_derotate(array, angle_list)
    mod.cube_derotate = derot
    assert _run(mod, "test_cube_derotate", select=lambda a: a == ("vip-fft", None, "constant", None)) == 1


def test_rescaling_known_answers():
    """tests/pre_3_10/test_preproc_rescaling.py:93-150: ``check_scal_vector`` ([2,8,4] -> [1,4,2], idempotent,
    TypeError) with the product's host function; x(1..10) rescaling and its inverse with the oracle's vip-fft
    restatement (the product rescales with GEMM operators on the GPU, pinned to the oracle on the GPU)."""
    import vip_b200.preproc.rescaling as r
    mod = _load("pre_3_10/test_preproc_rescaling.py")
    mod.check_scal_vector = r.check_scal_vector
    _run(mod, "test_check_scal_vector")
    ref_sdi = mod._find_indices_sdi
    mod._find_indices_sdi = r._find_indices_sdi
    assert _run(mod, "test_find_indices_sdi") == 7            # :152-169, seven known-answer index lists
    # beyond the reference's vectors: random wavelength grids, separations and `nframes` windows, same lists / errors
    rng = np.random.default_rng(0)
    for _ in range(300):
        z = int(rng.integers(3, 40))
        wl = np.sort(rng.uniform(0.9, 2.4, z))
        args = (wl.max() / wl, float(rng.uniform(2, 120)), int(rng.integers(0, z)), float(rng.uniform(2, 6)))
        kw = {"delta_sep": float(rng.choice([0.1, 0.5, 1.0])), "nframes": rng.choice([None, 2, 4, 8])}
        try:
            want = ref_sdi(*args, **kw)
        except RuntimeError:
            with pytest.raises(RuntimeError):
                r._find_indices_sdi(*args, **kw)
            continue
        np.testing.assert_array_equal(r._find_indices_sdi(*args, **kw), want)

    def rescale(cube, scal_list, full_output=True, inverse=False, y_in=None, x_in=None, imlib="vip-fft",
                interpolation=None, **kw):
        return O.cube_rescaling_wavelengths(cube, scal_list, full_output=full_output, inverse=inverse,
                                            y_in=y_in, x_in=x_in)
    mod.cube_rescaling_wavelengths = rescale
    _run(mod, "test_cube_rescaling_wavelengths", select=lambda a: a[0] == "vip-fft")


def test_svd_wrapper_reference_test_through_kernel_standins(monkeypatch):
    """tests/pre_3_10/test_pca_svd.py:10-21 (``U S V`` of the lapack mode reconstructs a 20 x 100 Gaussian matrix)
    run verbatim against ``vip_b200.psfsub.svd_wrapper``: the host logic (mode dispatch, ``full_output`` orientation
    of the lapack mode, dtype handling) is the product's, the kernels are the CPU stand-ins of
    tests/kernel_double.py (their GPU parity: tests/test_gpu_parity.py)."""
    import kernel_double
    vb = kernel_double.install(monkeypatch)
    mod = _load("pre_3_10/test_pca_svd.py")
    mod.svd_wrapper = vb.psfsub.svd_wrapper
    _run(mod, "test_svd_recons")
