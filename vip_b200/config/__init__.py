"""Parameter plumbing shared by the drop-in entry points (mirrors ``vip_hci.config``)."""
from .paramenum import *          # noqa: F401,F403
from .utils_param import separate_kwargs_dict, setup_parameters   # noqa: F401
from .utils_conf import check_array   # noqa: F401
