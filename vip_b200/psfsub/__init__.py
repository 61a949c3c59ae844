"""PSF-subtraction entry points (mirrors ``vip_hci.psfsub`` for the PCA hot path)."""
from .pca_fullfr import pca, PCA_Params                     # noqa: F401
from .pca_local import pca_annular, PCA_ANNULAR_Params      # noqa: F401
from .svd import svd_wrapper                                # noqa: F401
from .medsub import median_sub, MEDIAN_SUB_Params           # noqa: F401
from .incremental import pca_incremental                   # noqa: F401
