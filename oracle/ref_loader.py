"""Import the UNMODIFIED reference (vip_hci) from /root/reference for oracle pinning.

TEST INFRASTRUCTURE.  Only usable in the build container (``/root/reference`` does not
exist on the GPU box); used by ``tools/make_golden.py`` and
``tests/test_oracle_vs_reference.py`` (skipped when the checkout is absent).

vip_hci imports astropy / scikit-image / photutils / matplotlib / hciplot / ... at module
scope, none of which is installed here and none of which is *called* on the default
``pca`` / ``pca_annular`` / ``cube_derotate`` path.  A ``sys.meta_path`` finder serves
empty stand-ins for exactly the top-level packages that are missing.

One stand-in is not empty: ``skimage.draw.disk``, which ``var/shapes.py:88`` (``mask_circle``) calls on the
``mask_center_px`` / ``radius_int`` paths.  scikit-image is not installed, so the loader serves scikit-image's
published rule (``skimage/draw/draw.py`` ``ellipse`` -> ``_ellipse_in_shape``: bounding box
ceil(c - R) .. floor(c + R) clipped to ``shape``, pixels with ((r-cy)/R)^2 + ((c-cx)/R)^2 < 1).  The resulting
index sets are checked against the reference's own ``test_mask_circle`` known answers
(``tests/test_reference_own_tests.py``).  Likewise ``photutils.aperture.CircularAperture`` /
``aperture_photometry(method='exact')`` (``metrics/snr_source.py:393-397``) are served from the oracle's restatement
of the exact circle-pixel overlap, so that the reference's own ``snr`` / ``snrmap`` can run.
"""
import importlib.abc
import importlib.machinery
import importlib.util
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
# where the unmodified reference is looked for: the read-only checkout of the build container, else the offline pip
# install under baseline/_ref (git-ignored, shipped to the GPU box with the snapshot; bench.py's CPU arm uses it)
_CANDIDATES = ["/root/reference/src", os.path.join(os.path.dirname(_HERE), "baseline", "_ref")]
REFERENCE_SRC = os.environ.get("VIP_REFERENCE_SRC") or next(
    (c for c in _CANDIDATES if os.path.isdir(os.path.join(c, "vip_hci"))), _CANDIDATES[0])
_OPTIONAL = ["astropy", "skimage", "photutils", "matplotlib", "hciplot", "emcee", "nestle",
             "corner", "dataclass_builder", "pyds9", "munch", "ultranest"]


def _disk(center, radius, *, shape=None):
    """scikit-image ``draw.disk`` (see the module docstring)."""
    import numpy as np
    center = np.array(center, dtype=float)
    radii = np.array([radius, radius], dtype=float)
    upper_left = np.ceil(center - radii).astype(int)
    lower_right = np.floor(center + radii).astype(int)
    if shape is not None:
        upper_left = np.maximum(upper_left, np.array([0, 0]))
        lower_right = np.minimum(lower_right, np.array(shape[:2]) - 1)
    shifted = center - upper_left
    bshape = lower_right - upper_left + 1
    r_lim, c_lim = np.ogrid[0:float(bshape[0]), 0:float(bshape[1])]
    r, c = (r_lim - shifted[0]), (c_lim - shifted[1])
    rr, cc = np.nonzero((r / radius) ** 2 + (c / radius) ** 2 < 1)
    return rr + upper_left[0], cc + upper_left[1]


class _CircularAperture:
    """Stand-in for ``photutils.aperture.CircularAperture`` (positions as (x, y) pairs, radius r)."""

    def __init__(self, positions, r):
        import numpy as np
        pos = np.asarray(positions if isinstance(positions, np.ndarray) else list(positions), dtype=float)
        self.scalar = pos.ndim == 1                      # photutils also takes ONE (x, y) pair (frame_report)
        self.positions = [tuple(p) for p in np.atleast_2d(pos)]
        self.r = float(r)


def _aperture_photometry(data, apertures, method="exact", **kwargs):
    """Stand-in for ``photutils.aperture.aperture_photometry(..., method='exact')``: photutils is not installed, so
    the exact circle-pixel overlap sums come from the oracle's restatement (``vip_oracle.aperture_sums_exact``,
    pinned by photutils' documented known answer).  Lets the reference's own ``snr`` / ``snrmap`` run, which pins
    everything AROUND the aperture sums (centres, statistics, masks)."""
    if method != "exact":
        raise ImportError("photutils stand-in: only method='exact' is available")
    import numpy as np
    from . import vip_oracle
    xs = [p[0] for p in apertures.positions]
    ys = [p[1] for p in apertures.positions]
    return {"aperture_sum": vip_oracle.aperture_sums_exact(np.asarray(data), xs, ys, apertures.r)}


class _Stub(types.ModuleType):
    __path__ = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        modname = self.__name__
        if modname == "skimage.draw" and name == "disk":
            return _disk
        if modname == "photutils.aperture" and name == "CircularAperture":
            return _CircularAperture
        if modname == "photutils.aperture" and name == "aperture_photometry":
            return _aperture_photometry

        class _Missing(Warning):
            def __init__(self, *a, **k):
                raise ImportError(f"{modname}.{name} is not installed (stub)")
        _Missing.__name__ = name
        return _Missing


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def __init__(self, names):
        self.names = set(names)

    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in self.names:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        return _Stub(spec.name)

    def exec_module(self, module):
        pass


def available():
    return os.path.isdir(os.path.join(REFERENCE_SRC, "vip_hci"))


_loaded = None


def load():
    """Return the reference ``vip_hci`` package (imported once)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise ImportError(f"reference checkout not found under {REFERENCE_SRC}")
    missing = []
    for m in _OPTIONAL:
        try:
            if importlib.util.find_spec(m) is None:
                missing.append(m)
        except (ImportError, ValueError):
            missing.append(m)
    sys.meta_path.append(_Finder(missing))
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import vip_hci  # noqa
        import vip_hci.psfsub  # noqa
        import vip_hci.preproc  # noqa
        import vip_hci.var  # noqa
    _loaded = vip_hci
    return vip_hci
