"""Parallactic-angle vector sanitising (``vip_hci/preproc/parangles.py:405-458``)."""
import numpy as np


def check_pa_vector(angle_list, unit="deg"):
    """Return a copy in degrees with (1) negative angles shifted by +360 and (2), if any pair of
    consecutive angles is more than 180 deg apart, every angle below 180 shifted by +360."""
    if unit not in ("deg", "rad"):
        raise ValueError("The input unit should either be 'deg' or 'rad'")
    pa = np.array(angle_list, copy=True)
    if unit == "rad":
        pa = np.rad2deg(pa)
    pa[pa < 0] += 360
    if pa.shape[0] > 1 and np.any(np.abs(pa[1:] - pa[:-1]) > 180):
        pa[pa < 180] += 360
    return pa
