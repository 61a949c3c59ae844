"""FITS image I/O without astropy (drop-ins for ``vip_hci.fits.open_fits`` / ``write_fits`` / ``info_fits``)."""
from .fits import open_fits, write_fits, info_fits, verify_fits, byteswap_array, open_fits_device    # noqa: F401
