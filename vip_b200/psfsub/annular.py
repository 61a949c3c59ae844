"""Annular PCA on the B200 (drop-in for ``vip_hci.psfsub.pca_annular``) -- under construction."""
from dataclasses import dataclass
from enum import Enum
from typing import List, Tuple, Union

import numpy as np

from ..config.paramenum import Collapse, Imlib, Interpolation, SvdMode


@dataclass
class PCA_ANNULAR_Params:
    """Parameters of ``pca_annular`` in the reference's declaration order (``pca_local.py:43-70``)."""

    cube: np.ndarray = None
    angle_list: np.ndarray = None
    cube_ref: np.ndarray = None
    scale_list: np.ndarray = None
    radius_int: int = 0
    fwhm: float = 4
    asize: float = 4
    n_segments: Union[int, List[int], str] = 1
    delta_rot: Union[float, Tuple[float], List[float]] = (0.1, 1)
    delta_sep: Union[float, Tuple[float], List[float]] = (0.1, 1)
    ncomp: Union[int, Tuple, np.ndarray, str] = 1
    svd_mode: Enum = SvdMode.LAPACK
    nproc: int = 1
    min_frames_lib: int = 2
    max_frames_lib: int = 200
    tol: float = 1e-1
    scaling: Enum = None
    imlib: Enum = Imlib.VIPFFT
    interpolation: Enum = Interpolation.LANCZOS4
    collapse: Enum = Collapse.MEDIAN
    collapse_ifs: Enum = Collapse.MEAN
    ifs_collapse_range: Union[str, Tuple[int]] = "all"
    theta_init: int = 0
    weights: np.ndarray = None
    cube_sig: np.ndarray = None
    full_output: bool = False
    verbose: bool = True
    left_eigv: bool = False


def pca_annular(*all_args, **all_kwargs):
    raise NotImplementedError("vip_b200.pca_annular: GPU path under construction")
