"""FITS image I/O of ``vip_b200.fits`` (drop-ins for ``vip_hci.fits``: ``fits/fits.py:23-285``; astropy is not in this
image, the format is implemented from the FITS standard): round trips, header cards, scaled integers,
multi-extension files, memmap views, DATASUM verification, and the two FITS files the reference ships with its tests
(written by astropy) with the known answers its own tests rely on."""
import os

import numpy as np
import pytest

from vip_b200.fits import fits as F

REF_RES = "/root/reference/tests/pre_3_10"


def test_round_trip_types_header_and_checksum(tmp_path):
    rng = np.random.default_rng(0)
    cube = rng.normal(size=(3, 5, 7)) * 1e3
    name = str(tmp_path / "cube")
    hdr = {"OBJECT": "beta Pic b's host", "EXPTIME": 1.5, "NFRAMES": 3, "FLAG": True, "COMMENT": ["two", "lines"]}
    F.write_fits(name, cube, header=hdr, verbose=False)                      # '.fits' appended, float32 by default
    data, h = F.open_fits(name + ".fits", header=True, verbose=False)
    assert data.dtype == np.float32 and data.flags["C_CONTIGUOUS"]
    np.testing.assert_array_equal(data, cube.astype(np.float32))
    assert h["NAXIS"] == 3 and (h["NAXIS1"], h["NAXIS2"], h["NAXIS3"]) == (7, 5, 3) and h["BITPIX"] == -32
    assert h["OBJECT"] == "beta Pic b's host" and h["EXPTIME"] == 1.5 and h["NFRAMES"] == 3 and h["FLAG"] is True
    assert h["COMMENT"] == [" two", " lines"] or [c.strip() for c in h["COMMENT"]] == ["two", "lines"]
    assert os.path.getsize(name + ".fits") % 2880 == 0
    assert F.verify_fits(name + ".fits")
    np.testing.assert_array_equal(F.open_fits(name, verbose=False), data)    # name without the extension
    # float64 on request, both ways
    F.write_fits(name, cube, precision=np.float64, verbose=False)
    np.testing.assert_array_equal(F.open_fits(name, precision=np.float64, verbose=False), cube)
    # corrupt one data byte: DATASUM must notice
    raw = bytearray(open(name + ".fits", "rb").read())
    raw[-2880 + 5] ^= 0x40
    open(name + ".fits", "wb").write(bytes(raw))
    assert not F.verify_fits(name + ".fits")


def test_multi_extension_memmap_and_info(tmp_path, capsys):
    rng = np.random.default_rng(1)
    a, b = rng.normal(size=(4, 6, 6)).astype(np.float32), rng.normal(size=(9,)).astype(np.float32)
    name = str(tmp_path / "mef.fits")
    F.write_fits(name, (a, b), header=({"EXTNAME": "SCI"}, {"EXTNAME": "ANGLES"}), verbose=False)
    both, heads = F.open_fits(name, n=-2, header=True, verbose=False)
    np.testing.assert_array_equal(both[0], a)
    np.testing.assert_array_equal(both[1], b)
    assert heads[1]["XTENSION"] == "IMAGE" and heads[1]["EXTNAME"] == "ANGLES"
    np.testing.assert_array_equal(F.open_fits(name, n=1, verbose=False), b)
    mm = F.open_fits(name, return_memmap=True, verbose=False)
    assert isinstance(mm, np.memmap) and mm.dtype == np.dtype(">f4") and mm.shape == a.shape
    np.testing.assert_array_equal(F.byteswap_array(np.asarray(mm[1:3])), a[1:3])      # the slices `batch` reads
    F.info_fits(name)
    assert "ImageHDU" in capsys.readouterr().out
    with pytest.raises(ValueError):
        F.write_fits(name, (a, b), header=({}, {}, {}), verbose=False)


def test_scaled_integers_and_foreign_cards(tmp_path):
    """A file as another writer would produce it: 16-bit samples with BZERO = 32768 (the unsigned convention), BSCALE,
    a string with doubled quotes, a D exponent, HISTORY, a blank card, an unknown extension to skip, an IMAGE after it."""
    vals = np.array([[0, 1, 2], [65535, 40000, 7]], dtype=np.uint16)
    stored = (vals.astype(np.int32) - 32768).astype(">i2")

    def block(cards, data=b""):
        head = "".join(f"{c:<80s}" for c in cards + ["END"]).encode()
        head += b" " * (-len(head) % 2880)
        return head + data + b"\0" * (-len(data) % 2880)

    primary = block(["SIMPLE  =                    T", "BITPIX  =                   16", "NAXIS   =                    2",
                     "NAXIS1  =                    3", "NAXIS2  =                    2", "EXTEND  =                    T",
                     "BZERO   =                32768", "BSCALE  =                  2.0",
                     "OBSERVER= 'O''Brien '           / who", "GAIN    =               1.25D1 / e-/ADU",
                     "HISTORY made by hand", ""], stored.tobytes())
    table = block(["XTENSION= 'BINTABLE'", "BITPIX  =                    8", "NAXIS   =                    2",
                   "NAXIS1  =                   10", "NAXIS2  =                    3", "PCOUNT  =                    0",
                   "GCOUNT  =                    1", "TFIELDS =                    1"], b"x" * 30)
    image = block(["XTENSION= 'IMAGE   '", "BITPIX  =                  -64", "NAXIS   =                    1",
                   "NAXIS1  =                    2", "PCOUNT  =                    0", "GCOUNT  =                    1"],
                  np.array([1.5, -2.25], dtype=">f8").tobytes())
    name = str(tmp_path / "foreign.fits")
    open(name, "wb").write(primary + table + image)
    data, h = F.open_fits(name, header=True, verbose=False)
    np.testing.assert_array_equal(data, ((vals.astype(np.float64) - 32768) * 2.0 + 32768).astype(np.float32))
    assert h["OBSERVER"] == "O'Brien" and h.comments["OBSERVER"] == "who" and h["GAIN"] == 12.5
    assert h["HISTORY"] == [" made by hand"] or h["HISTORY"][0].strip() == "made by hand"
    assert F.open_fits(name, n=1, verbose=False) is None                       # not an image: skipped by its size
    np.testing.assert_array_equal(F.open_fits(name, n=2, precision=np.float64, verbose=False), [1.5, -2.25])
    with pytest.raises(OSError):
        open(str(tmp_path / "noend.fits"), "wb").write(primary[:800])
        F.open_fits(str(tmp_path / "noend.fits"), verbose=False)


@pytest.mark.skipif(not os.path.isdir(REF_RES), reason="reference checkout not present")
def test_files_shipped_with_the_reference():
    """``SPHERE_satspots_centered.fits`` (151 x 151, written by astropy): the reference's own recentering test places
    the four satellite spots at (x, y) = (41, 109), (109, 109), (41, 41), (109, 41) of this CENTERED frame
    (``tests/pre_3_10/test_preproc_recentering.py:585``); ``naco_betapic_single.fits`` (101 x 101) is used with
    ``center_fr1=(51, 51)`` (:541).  A reader that mis-parses the header, the byte order or the axis order cannot
    reproduce either."""
    img = F.open_fits(os.path.join(REF_RES, "SPHERE_satspots_centered.fits"), verbose=False)
    assert img.shape == (151, 151) and img.dtype == np.float32 and np.isfinite(img).all()
    halo = np.median(img)
    for x, y in ((41, 109), (109, 109), (41, 41), (109, 41)):
        box = img[y - 12:y + 13, x - 12:x + 13]
        by, bx = np.unravel_index(np.argmax(box), box.shape)
        assert abs(by - 12) <= 3 and abs(bx - 12) <= 3 and box.max() > 2 * halo, (x, y, by, bx)
    # centred: the frame is close to its own point reflection about (75, 75)
    assert np.corrcoef(img.ravel(), img[::-1, ::-1].ravel())[0, 1] > 0.8
    img = F.open_fits(os.path.join(REF_RES, "naco_betapic_single"), verbose=False)      # '.fits' appended
    assert img.shape == (101, 101) and img.dtype == np.float32
    # a coronagraphic NACO frame ("negative=True" in the reference's test): the star sits in the dark hole at (51, 51)
    yy, xx = np.mgrid[:101, :101]
    ring = (np.hypot(yy - 51, xx - 51) >= 2) & (np.hypot(yy - 51, xx - 51) < 4)
    core = np.hypot(yy - 51, xx - 51) < 2
    assert img[ring].mean() > 1.3 * img[core].mean()        # 2151 vs 1374
    far = np.hypot(yy - 51, xx - 51) > 30
    assert img[far].mean() < 0.05 * img[ring].mean()        # the halo falls off around (51, 51)
