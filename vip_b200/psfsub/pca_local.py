"""Annular PCA (drop-in for ``vip_hci.psfsub.pca_annular``) -- filled in by annular.py."""
from .annular import pca_annular, PCA_ANNULAR_Params    # noqa: F401
