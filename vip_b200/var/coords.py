"""Centre / distance conventions (``vip_hci/var/coords.py:21-24, 61-100``)."""
import numpy as np


def frame_center(array, verbose=False):
    """(cy, cx) of a frame / cube / 4-d cube (or of a shape tuple): N/2 for even N, (N-1)/2 for odd N."""
    shape = array.shape if hasattr(array, "shape") else tuple(array)
    if len(shape) not in (2, 3, 4):
        raise ValueError("`array` is not a 2d, 3d or 4d array")
    ny, nx = shape[-2], shape[-1]
    cy = ny // 2 if ny % 2 == 0 else (ny - 1) // 2
    cx = nx // 2 if nx % 2 == 0 else (nx - 1) // 2
    if verbose:
        print("Center px coordinates at x,y = ({}, {})".format(cx, cy))
    return int(cy), int(cx)


def dist(yc, xc, y1, x1):
    """Euclidean distance between two points."""
    return np.sqrt((yc - y1) ** 2 + (xc - x1) ** 2)
