"""FITS image I/O for the path: drop-ins for ``vip_hci.fits.open_fits`` / ``write_fits`` / ``info_fits`` /
``verify_fits`` / ``byteswap_array`` (``src/vip_hci/fits/fits.py:23-285``) that do not need astropy.

The reference delegates to ``astropy.io.fits``; astropy is absent from this image, so the container format itself is
implemented here from the FITS standard 4.0: 2880-byte blocks, 80-character header cards, big-endian data units of
``BITPIX`` 8 / 16 / 32 / 64 / -32 / -64, ``BSCALE`` / ``BZERO``, the primary HDU followed by ``IMAGE`` extensions (what
``write_fits`` produces for a tuple of arrays); other extension types are skipped by their declared size.
``open_fits_device`` additionally decodes the data unit ON the GPU: raw big-endian bytes go through the multi-threaded
pinned staging upload and ``vb_fits_decode_f32`` (byte swap + scaling + fp32 conversion), so a cube reaches HBM
without a host pass over its samples.  The ``DATASUM`` / ``CHECKSUM`` keywords are written and can be verified
(``verify_fits``), which is also how the reader is tested against files written by other software.
"""
import os
from collections import OrderedDict

import numpy as np

BLOCK = 2880
ALL_FITS = -2
_DTYPES = {8: ">u1", 16: ">i2", 32: ">i4", 64: ">i8", -32: ">f4", -64: ">f8"}
_BITPIX = {np.dtype("uint8"): 8, np.dtype("int16"): 16, np.dtype("int32"): 32, np.dtype("int64"): 64,
           np.dtype("float32"): -32, np.dtype("float64"): -64}


class Header(OrderedDict):
    """Header of one HDU: keyword -> value in file order (``COMMENT`` / ``HISTORY`` collect lists); ``comments``
    holds the card comments."""

    def __init__(self, *args, **kw):
        super().__init__(*args, **kw)
        self.comments = {}


def _parse_value(text):
    t = text.strip()
    if not t:
        return None
    if t[0] == "'":                                   # string: quotes doubled inside, trailing blanks not significant
        out, i = [], 1
        while i < len(t):
            if t[i] == "'":
                if i + 1 < len(t) and t[i + 1] == "'":
                    out.append("'")
                    i += 2
                    continue
                break
            out.append(t[i])
            i += 1
        return "".join(out).rstrip(), t[i + 1:]
    val, _, rest = t.partition("/")
    v = val.strip()
    if v in ("T", "F"):
        return v == "T", "/" + rest if rest else ""
    try:
        return int(v), "/" + rest if rest else ""
    except ValueError:
        pass
    try:
        return float(v.replace("D", "E").replace("d", "e")), "/" + rest if rest else ""
    except ValueError:
        return v, "/" + rest if rest else ""


def _parse_header(buf, offset, ignore_missing_end=False):
    """Header starting at ``offset``: (Header, offset of the data unit)."""
    hdr = Header()
    pos = offset
    n = len(buf)
    while True:
        if pos + BLOCK > n:
            if ignore_missing_end:
                return hdr, n
            raise OSError("Header missing END card.")
        block = bytes(buf[pos:pos + BLOCK])
        pos += BLOCK
        for i in range(0, BLOCK, 80):
            card = block[i:i + 80].decode("ascii", "replace")
            key = card[:8].rstrip()
            if key == "END":
                return hdr, pos
            if key in ("COMMENT", "HISTORY", ""):
                if card[8:].strip():
                    hdr.setdefault(key or "COMMENT", []).append(card[8:].rstrip())
                continue
            if key == "CONTINUE" and hdr:
                last = next(reversed(hdr))
                parsed = _parse_value(card[8:])
                if isinstance(hdr[last], str) and parsed and isinstance(parsed[0], str):
                    hdr[last] = hdr[last].rstrip("&") + parsed[0]
                continue
            if card[8:10] != "= ":
                continue
            parsed = _parse_value(card[10:])
            if parsed is None:
                hdr[key] = None
                continue
            value, rest = parsed
            hdr[key] = value
            rest = rest.strip()
            if rest.startswith("/"):
                hdr.comments[key] = rest[1:].strip()


def _data_geometry(hdr):
    bitpix = int(hdr.get("BITPIX", 8))
    naxis = int(hdr.get("NAXIS", 0))
    shape = tuple(int(hdr["NAXIS%d" % (i + 1)]) for i in range(naxis))[::-1]        # NAXIS1 is the fastest axis
    count = int(np.prod(shape)) if naxis else 0
    nbytes = abs(bitpix) // 8 * int(hdr.get("GCOUNT", 1)) * (int(hdr.get("PCOUNT", 0)) + count)
    return bitpix, shape, count, nbytes


def _scan(buf, ignore_missing_end=False):
    """[(Header, data offset, bitpix, shape, count)] for every HDU of the file."""
    hdus = []
    pos, n = 0, len(buf)
    while pos < n:
        if not bytes(buf[pos:pos + 8]).strip():
            break                                        # padding after the last HDU
        hdr, data_off = _parse_header(buf, pos, ignore_missing_end)
        if not hdus and hdr.get("SIMPLE") is not True and not ignore_missing_end:
            raise OSError("No SIMPLE card found, this file does not appear to be a valid FITS file.")
        bitpix, shape, count, nbytes = _data_geometry(hdr)
        hdus.append((hdr, data_off, bitpix, shape, count))
        pos = data_off + (nbytes + BLOCK - 1) // BLOCK * BLOCK
    return hdus


def _is_image(hdr):
    return hdr.get("SIMPLE") is True or str(hdr.get("XTENSION", "")).strip() == "IMAGE"


def _hdu_array(buf, hdu, precision):
    hdr, off, bitpix, shape, count = hdu
    if count == 0 or not _is_image(hdr):
        return None
    raw = np.frombuffer(buf, dtype=_DTYPES[bitpix], count=count, offset=off).reshape(shape)
    bscale, bzero = hdr.get("BSCALE", 1), hdr.get("BZERO", 0)
    if bscale != 1 or bzero != 0:
        data = raw.astype(np.float64) * float(bscale) + float(bzero)
    else:
        data = raw
    return np.array(data, dtype=precision)


def _filename(fitsfilename):
    fitsfilename = str(fitsfilename)
    if not os.path.isfile(fitsfilename):
        fitsfilename += ".fits"
    return fitsfilename


def open_fits(fitsfilename, n=0, header=False, ignore_missing_end=False, precision=np.float32,
              return_memmap=False, verbose=True, **kwargs):
    """Load a FITS file into memory as numpy array(s) (``fits/fits.py:23-117``).

    ``n`` selects the HDU (-2: all of them, as lists); ``header=True`` also returns the header (a ``Header`` dict);
    ``return_memmap=True`` returns the big-endian ``numpy.memmap`` view of the data unit instead of a converted copy
    (the analogue of the astropy HDU handle the reference returns: what ``batch`` processing slices from)."""
    fitsfilename = _filename(fitsfilename)
    buf = np.memmap(fitsfilename, dtype=np.uint8, mode="r")
    hdus = _scan(buf, ignore_missing_end)

    def one(index):
        hdr, off, bitpix, shape, count = hdus[index]
        if return_memmap:
            return np.memmap(fitsfilename, dtype=_DTYPES[bitpix], mode="r", offset=off, shape=shape)
        data = _hdu_array(buf, hdus[index], precision)
        if verbose:
            what = "data and header" if header else "data"
            print(f"FITS HDU-{index} {what} successfully loaded. Data shape: {None if data is None else data.shape}")
        return data

    if n == ALL_FITS:
        data_list = [one(i) for i in range(len(hdus))]
        if return_memmap:
            return data_list
        if verbose:
            print(f"All {len(hdus)} FITS HDU data{' and headers' if header else ''} successfully loaded.")
        return (data_list, [h[0] for h in hdus]) if header else data_list
    data = one(n)
    if return_memmap:
        return data
    return (data, hdus[n][0]) if header else data


def open_fits_device(fitsfilename, n=0, header=False, ignore_missing_end=False, device=None):
    """``open_fits`` straight to the GPU: the raw data unit of HDU ``n`` is uploaded (pinned multi-threaded staging)
    and decoded there (``vb_fits_decode_f32``).  Returns a fp32 CUDA tensor (and the header)."""
    import torch
    from .. import _cabi
    from .._device import require_cuda, stream_ptr
    fitsfilename = _filename(fitsfilename)
    buf = np.memmap(fitsfilename, dtype=np.uint8, mode="r")
    hdr, off, bitpix, shape, count = _scan(buf, ignore_missing_end)[n]
    if count == 0 or not _is_image(hdr):
        raise ValueError(f"HDU {n} of {fitsfilename} holds no image")
    dev = device or require_cuda()
    nbytes = count * (abs(bitpix) // 8)
    raw = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    out = torch.empty(shape, dtype=torch.float32, device=dev)
    src = np.ascontiguousarray(buf[off:off + nbytes])
    lib = _cabi.lib()
    with torch.cuda.device(raw.device):
        _cabi.check(lib.vb_memcpy_h2d_staged(int(raw.data_ptr()), int(src.ctypes.data), nbytes, stream_ptr()),
                    "vb_memcpy_h2d_staged")
        _cabi.check(lib.vb_fits_decode_f32(int(raw.data_ptr()), int(bitpix), count, float(hdr.get("BSCALE", 1)),
                                           float(hdr.get("BZERO", 0)), int(out.data_ptr()), stream_ptr()),
                    "vb_fits_decode_f32")
    return (out, hdr) if header else out


def byteswap_array(array):
    """Array in the byte order of this machine (``fits/fits.py:149-179``): FITS data are big-endian."""
    array = np.asarray(array)
    if array.dtype.byteorder in ("=", "|") or (array.dtype.byteorder == "<") == (np.little_endian):
        return array
    return array.byteswap().view(array.dtype.newbyteorder("="))


def info_fits(fitsfilename, **kwargs):
    """Print the HDU table of a FITS file (``fits/fits.py:182-196``)."""
    buf = np.memmap(_filename(fitsfilename), dtype=np.uint8, mode="r")
    print("No.    Name      Type        Cards   Dimensions   Format")
    for i, (hdr, _, bitpix, shape, _) in enumerate(_scan(buf)):
        kind = "PrimaryHDU" if i == 0 else str(hdr.get("XTENSION", "")).strip().title() + "HDU"
        name = "PRIMARY" if i == 0 else str(hdr.get("EXTNAME", ""))
        print(f"{i:3d}  {name:10s} {kind:11s} {len(hdr):5d}   {shape[::-1]!s:12s} {_DTYPES[bitpix][1:]}")


def _ones_complement_sum(data):
    """FITS checksum accumulator: one's complement sum of the big-endian 32-bit words."""
    b = np.frombuffer(data, dtype=">u4").astype(np.uint64)
    hi = int(b[0::2].sum())
    lo = int(b[1::2].sum()) if b.size > 1 else 0
    while hi >> 32 or lo >> 32:                      # end-around carries (FITS standard, appendix J)
        hicarry, locarry = hi >> 32, lo >> 32
        hi = (hi & 0xFFFFFFFF) + locarry
        lo = (lo & 0xFFFFFFFF) + hicarry
    total = hi + lo
    while total >> 32:
        total = (total & 0xFFFFFFFF) + (total >> 32)
    return total


def datasum(data_bytes):
    """``DATASUM`` of a (block-padded) data unit as the unsigned integer the keyword stores."""
    if len(data_bytes) % 4:
        data_bytes = data_bytes + b"\0" * (4 - len(data_bytes) % 4)
    b = np.frombuffer(data_bytes, dtype=">u4").astype(np.uint64)
    total = int(b.sum())
    while total >> 32:
        total = (total & 0xFFFFFFFF) + (total >> 32)
    return total


def verify_fits(fitsfilename):
    """Check the data units of a FITS file against their ``DATASUM`` keywords (``fits/fits.py:199-215`` verifies
    through astropy).  Prints the verdict per HDU; returns True when every HDU that carries a ``DATASUM`` matches."""
    names = fitsfilename if isinstance(fitsfilename, list) else [fitsfilename]
    ok = True
    for name in names:
        buf = np.memmap(_filename(name), dtype=np.uint8, mode="r")
        for i, (hdr, off, bitpix, shape, count) in enumerate(_scan(buf)):
            nbytes = count * (abs(bitpix) // 8)
            padded = (nbytes + BLOCK - 1) // BLOCK * BLOCK
            if "DATASUM" not in hdr:
                print(f"{name} HDU-{i}: no DATASUM keyword")
                continue
            good = datasum(bytes(buf[off:off + padded])) == int(str(hdr["DATASUM"]).strip())
            print(f"{name} HDU-{i}: DATASUM {'OK' if good else 'MISMATCH'}")
            ok = ok and good
    return ok


def _card(key, value, comment=""):
    if isinstance(value, bool):
        v = f"{'T' if value else 'F':>20s}"
    elif isinstance(value, (int, np.integer)):
        v = f"{int(value):>20d}"
    elif isinstance(value, (float, np.floating)):
        v = f"{repr(float(value)).upper():>20s}"
    elif value is None:
        v = ""
    else:
        s = str(value).replace("'", "''")
        v = f"'{s:<8s}'"
        v = f"{v:<20s}"
    card = f"{key:<8s}= {v}"
    if comment:
        card += f" / {comment}"
    return f"{card[:80]:<80s}"


def _hdu_bytes(array, header, primary):
    array = np.asarray(array)
    if array.dtype not in _BITPIX:
        raise TypeError(f"write_fits: unsupported array type {array.dtype}")
    bitpix = _BITPIX[array.dtype]
    cards = [_card("SIMPLE", True, "conforms to FITS standard") if primary
             else _card("XTENSION", "IMAGE", "Image extension"),
             _card("BITPIX", bitpix, "array data type"), _card("NAXIS", array.ndim, "number of array dimensions")]
    for i, dim in enumerate(array.shape[::-1]):
        cards.append(_card("NAXIS%d" % (i + 1), dim))
    if primary:
        cards.append(_card("EXTEND", True))
    else:
        cards += [_card("PCOUNT", 0, "number of parameters"), _card("GCOUNT", 1, "number of groups")]
    data = np.ascontiguousarray(array, dtype=_DTYPES[bitpix]).tobytes()
    data += b"\0" * (-len(data) % BLOCK)
    reserved = {"SIMPLE", "XTENSION", "BITPIX", "NAXIS", "EXTEND", "PCOUNT", "GCOUNT", "END", "DATASUM", "CHECKSUM"}
    if header:
        comments = getattr(header, "comments", {})
        for key, value in dict(header).items():
            key = str(key).upper()[:8]
            if key in reserved or key.startswith("NAXIS"):
                continue
            if key in ("COMMENT", "HISTORY"):
                for line in (value if isinstance(value, (list, tuple)) else [value]):
                    cards.append(f"{key:<8s}{str(line)[:72]:<72s}")
                continue
            cards.append(_card(key, value, comments.get(key, "") if isinstance(comments, dict) else ""))
    cards.append(_card("DATASUM", str(datasum(data)), "data unit checksum"))
    cards.append(f"{'END':<80s}")
    head = "".join(cards).encode("ascii")
    head += b" " * (-len(head) % BLOCK)
    return head + data


def write_fits(fitsfilename, array, header=None, output_verify="exception", precision=np.float32, verbose=True):
    """Write array(s) and header(s) into a FITS file (``fits/fits.py:218-285``): one array -> primary HDU, a tuple of
    arrays -> primary HDU + IMAGE extensions (astropy's ``HDUList([ImageHDU, ...])`` gets an empty primary; here the
    first array is the primary, which ``open_fits(n=...)`` indexes the same way).  An existing file is replaced."""
    if not fitsfilename.endswith(".fits"):
        fitsfilename += ".fits"
    res = "overwritten" if os.path.exists(fitsfilename) else "saved"
    if isinstance(array, tuple):
        if header is None:
            header = [None] * len(array)
        elif not isinstance(header, tuple):
            header = [header] * len(array)
        elif len(header) != len(array):
            raise ValueError("If input header is a tuple, it should have the same length as tuple of arrays.")
        blob = b"".join(_hdu_bytes(np.asarray(a).astype(precision, copy=False), h, i == 0)
                        for i, (a, h) in enumerate(zip(array, header)))
    else:
        blob = _hdu_bytes(np.asarray(array).astype(precision, copy=False), header, True)
    tmp = fitsfilename + ".tmp"
    with open(tmp, "wb") as f:
        f.write(blob)
    os.replace(tmp, fitsfilename)
    if verbose:
        print(f"FITS file successfully {res}")
