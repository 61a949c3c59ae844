"""String-valued enums accepted by ``pca`` / ``pca_annular`` (``vip_hci/config/paramenum.py:8-177``).

Members are ``str`` subclasses, so plain strings (``"lapack"``) and members
(``SvdMode.LAPACK``) compare equal, as in the reference.
"""
from enum import Enum

ALGO_KEY = "algo_params"


class SvdMode(str, Enum):
    LAPACK = "lapack"
    ARPACK = "arpack"
    EIGEN = "eigen"
    RANDSVD = "randsvd"
    CUPY = "cupy"
    EIGENCUPY = "eigencupy"
    RANDCUPY = "randcupy"
    PYTORCH = "pytorch"
    EIGENPYTORCH = "eigenpytorch"
    RANDPYTORCH = "randpytorch"


class Scaling(str, Enum):
    TEMPMEAN = "temp-mean"
    SPATMEAN = "spat-mean"
    TEMPSTANDARD = "temp-standard"
    SPATSTANDARD = "spat-standard"


class Adimsdi(str, Enum):
    DOUBLE = "double"
    SINGLE = "single"
    SKIPADI = "skipadi"


class Imlib(str, Enum):
    OPENCV = "opencv"
    SKIMAGE = "skimage"
    NDIMAGE = "ndimage"
    VIPFFT = "vip-fft"


class Interpolation(str, Enum):
    NEARNEIG = "nearneig"
    BILINEAR = "bilinear"
    BIQUADRATIC = "biquadratic"
    BICUBIC = "bicubic"
    BIQUARTIC = "biquartic"
    BIQUINTIC = "biquintic"
    LANCZOS4 = "lanczos4"


class Collapse(str, Enum):
    MEDIAN = "median"
    MEAN = "mean"
    SUM = "sum"
    TRIMMEAN = "trimmean"


__all__ = ["ALGO_KEY", "SvdMode", "Scaling", "Adimsdi", "Imlib", "Interpolation", "Collapse"]
