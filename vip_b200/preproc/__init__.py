"""Image operations on the hot path (mirrors the used part of ``vip_hci.preproc``)."""
from .derotation import (cube_derotate, frame_rotate, _find_indices_adi, _compute_pa_thresh,   # noqa: F401
                         _define_annuli, rotation_geometry, rotation_scalars)
from .subsampling import cube_collapse          # noqa: F401
from .parangles import check_pa_vector          # noqa: F401
from .recentering import frame_shift, cube_shift    # noqa: F401
from .rescaling import check_scal_vector, _find_indices_sdi    # noqa: F401
