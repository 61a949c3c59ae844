"""Geometry helpers on the hot path (mirrors the used part of ``vip_hci.var``)."""
from .coords import frame_center, dist                      # noqa: F401
from .shapes import (get_annulus_segments, mask_circle, disk_indices, reshape_matrix)   # noqa: F401
