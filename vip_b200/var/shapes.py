"""Index sets that must be bit-exact w.r.t. the reference (computed on the host in fp64 with the
same numpy expressions): annulus segments and circular masks (``vip_hci/var/shapes.py``)."""
import numpy as np

from .coords import frame_center


def disk_indices(cy, cx, radius, shape):
    """(rows, cols) strictly inside the circle -- the index set ``skimage.draw.disk`` yields, which
    ``mask_circle`` relies on (``var/shapes.py:88``)."""
    rr, cc = np.ogrid[: shape[0], : shape[1]]
    return np.nonzero(((rr - cy) / radius) ** 2 + ((cc - cx) / radius) ** 2 < 1)


def circle_mask(shape, radius):
    """Boolean (H,W) mask of the pixels ``mask_circle(mode='in')`` overwrites for a 3-d cube.

    The reference indexes cubes as ``[:, ind[1], ind[0]]`` (``var/shapes.py:101``), i.e. with the
    disk indices swapped; that is reproduced here."""
    m = np.zeros(shape, dtype=bool)
    if radius == 0:
        return m
    cy, cx = frame_center(shape)
    ind = disk_indices(cy, cx, radius, shape)
    m[ind[1], ind[0]] = True
    return m


def mask_circle(array, radius, fillwith=0, mode="in", cy=None, cx=None, output="masked_arr"):
    """Copy of a numpy frame/cube with the pixels inside (``mode='in'``) or outside (``'out'``) the disk set to
    ``fillwith``, or the boolean keep-mask of the disk (``output='bool_mask'``) (``var/shapes.py:38-113``).
    Cubes are indexed with the disk indices swapped, as the reference does (:101, :109)."""
    if not isinstance(fillwith, (int, float)):
        raise ValueError("`fillwith` must be integer, float or np.nan")
    shape = (array.shape[-2], array.shape[-1])
    if cy is None or cx is None:
        cy, cx = frame_center(shape)
    if radius == 0:
        keep_all = mode == "in"
        if output == "bool_mask":
            return np.full(shape, keep_all, dtype=bool)
        return array * keep_all
    rows, cols = disk_indices(cy, cx, radius, shape)
    if output == "bool_mask":
        keep = np.ones(shape, dtype=bool)
        keep[rows, cols] = False
        return keep
    if output != "masked_arr":
        raise ValueError("`output` not recognized")
    # 2-d frames use (rows, cols); cubes use the swapped pair on their last two axes
    sel = (rows, cols) if array.ndim == 2 else (Ellipsis, cols, rows)
    if mode == "in":
        out = array.copy()
        out[sel] = fillwith
    elif mode == "out":
        out = np.full_like(array, fillwith)
        out[sel] = array[sel]
    else:
        raise ValueError("`mode` not recognized")
    return out


def get_annulus_segments(data, inner_radius, width, nsegm=1, theta_init=0):
    """List of ``(yy, xx)`` index arrays, one per azimuthal segment of the annulus
    ``inner_radius <= r < inner_radius + width`` (``var/shapes.py:474-581``, mode='ind')."""
    shape = data.shape if hasattr(data, "shape") else tuple(data)
    if not isinstance(nsegm, int):
        raise TypeError("`nsegm` must be an integer")
    ny, nx = shape[-2], shape[-1]
    cy, cx = frame_center((ny, nx))
    span = np.deg2rad(int(np.ceil(360 / nsegm)))
    two_pi = 2 * np.pi
    yy, xx = np.mgrid[:ny, :nx]
    rad = np.sqrt((xx - cx) ** 2 + (yy - cy) ** 2)
    phi = np.arctan2(yy - cy, xx - cx) % two_pi
    ring = (rad >= inner_radius) & (rad < inner_radius + width)
    segments = []
    for i in range(nsegm):
        lo = np.deg2rad(theta_init) + i * span
        hi = lo + span
        if lo < two_pi < hi:
            sel = ring & (phi >= lo) & (phi <= two_pi) | ring & (phi >= 0) & (phi < hi - two_pi)
        elif lo >= two_pi and hi > two_pi:
            sel = ring & (phi >= lo - two_pi) & (phi < hi - two_pi)
        else:
            sel = ring & (phi >= lo) & (phi < hi)
        segments.append(np.where(sel))
    return segments


def reshape_matrix(array, y, x):
    """(nframes, y*x) -> (nframes, y, x) (``var/shapes.py:876-910``)."""
    return array.reshape(array.shape[0], y, x)
