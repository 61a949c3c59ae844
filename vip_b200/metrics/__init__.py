"""S/N of test resolution elements and S/N maps on the GPU (drop-ins for ``vip_hci.metrics``)."""
from .snr_source import snr, snrmap, indep_ap_centers      # noqa: F401
from .detection import detection, peak_local_max, sigma_clipped_stats, fit_gaussian2d      # noqa: F401
