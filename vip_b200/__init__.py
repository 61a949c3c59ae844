"""vip_b200 -- Blackwell-native ADI/RDI PCA speckle-subtraction path behind the call signatures
of ``vip_hci.psfsub.pca``, ``vip_hci.psfsub.pca_annular`` and ``vip_hci.preproc.cube_derotate``.

Host orchestration is Python; all arithmetic runs in hand-written sm_100a CUDA kernels reached
through the C ABI of ``libvipb200.so`` (``include/vip_b200.h``).  No CPU fallback.
"""
__version__ = "0.1.0"

from . import config, fits, metrics, preproc, psfsub, var        # noqa: F401
from .psfsub import pca, pca_annular, median_sub                        # noqa: F401
from .preproc import cube_derotate, cube_collapse, cube_shift, frame_shift           # noqa: F401
from .metrics import snr, snrmap, detection                                             # noqa: F401
from .fits import open_fits, write_fits                                                 # noqa: F401
