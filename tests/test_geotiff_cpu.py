"""GeoTIFF reader / writer behind the C ABI (mb_tiff_*, csrc/tiff_io.cu): the golden fixture cut
from the reference's bundled TWI raster (original GDAL-compressed tiles), cross-checks against Pillow / libtiff in both
directions, round trips at ragged sizes, refusals.  No GPU involved."""
import os
import struct

import numpy as np
import pytest

import machisplin_b200 as mb
from machisplin_b200._lib import MbError

PIL = pytest.importorskip("PIL.Image")
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def same(a, b):
    return np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(np.nan_to_num(a), np.nan_to_num(b))


def test_golden_tiles_of_the_reference_raster():
    """INT16, 128 x 128 tiles, LZW exactly as GDAL wrote them into inst/extdata/TWI.tif, NoData -32768 -> NaN."""
    z = np.load(os.path.join(GOLD, "twi_2x2_tiles.npz"))
    info = mb.raster_info(os.path.join(GOLD, "twi_2x2_tiles.tif"))
    assert (info.dtype, info.compression, info.tiled, info.chunk, info.nodata) == ("INT2S", 5, True, (128, 128), -32768.0)
    g, a = mb.read_raster(os.path.join(GOLD, "twi_2x2_tiles.tif"))
    ref = z["values"].astype(np.float32)
    ref[z["values"] == z["nodata"]] = np.nan
    assert np.isnan(ref).sum() > 0
    assert same(a, ref)
    assert g.nrow == 256 and g.ncol == 256
    assert abs(g.xmin - float(z["xmin"])) < 1e-12 and abs(g.ymax - float(z["ymax"])) < 1e-12
    assert abs(g.rx - float(z["rx"])) < 1e-15 and abs(g.ry - float(z["ry"])) < 1e-15


@pytest.mark.skipif(not os.path.isdir("/root/reference/inst/extdata"), reason="reference checkout not present")
@pytest.mark.parametrize("name", ["alt", "slope", "TWI"])
def test_bundled_rasters_decode_like_libtiff(name):
    """the three rasters of README Example 1, whole files (build container only)"""
    p = f"/root/reference/inst/extdata/{name}.tif"
    info = mb.raster_info(p)
    g, a = mb.read_raster(p)
    ref = np.array(PIL.open(p)).astype(np.float32)
    ref[ref == np.float32(info.nodata)] = np.nan
    assert same(a, ref)
    gold = np.load(os.path.join(GOLD, "bundled_c1.npz"), allow_pickle=True)
    if "geom" in gold.files:
        assert np.allclose(np.asarray(gold["geom"], dtype=float)[:4], g.as_tuple()[:4], rtol=0, atol=1e-6)
    assert info.epsg == 4326 and info.georeferenced


@pytest.mark.parametrize("shape", [(1, 1), (256, 256), (300, 517), (700, 3), (257, 1025)])
@pytest.mark.parametrize("compression", ["NONE", "LZW"])
def test_write_read_round_trip(tmp_path, shape, compression):
    rng = np.random.default_rng(shape[0] * 7 + shape[1])
    a = rng.standard_normal(shape).astype(np.float32)
    a[rng.random(shape) < 0.05] = np.nan
    if shape[0] > 100:
        a[40:90] = 3.25                      # long runs: deep LZW strings
    geom = mb.Geom(-77.5, -77.5 + 0.001 * shape[1], -6.0, -6.0 + 0.002 * shape[0], shape[0], shape[1])
    p = str(tmp_path / "r.tif")
    mb.write_raster(p, geom, a, compression=compression, epsg=4326)
    info = mb.raster_info(p)
    assert info.dtype == "FLT4S" and info.tiled and info.chunk == (256, 256) and info.epsg == 4326
    assert np.allclose(info.geom.as_tuple()[:4], geom.as_tuple()[:4], rtol=0, atol=1e-12)
    g, b = mb.read_raster(p, threads=3)
    assert same(a, b)
    # libtiff reads what we wrote
    c = np.array(PIL.open(p))
    assert c.dtype == np.float32 and same(a, c)
    # float64 input is stored as FLT4S like terra::writeRaster's default
    mb.write_raster(p, geom, a.astype(np.float64), compression=compression)
    assert same(mb.read_raster(p)[1], a)


def test_lzw_table_resets_and_long_strings(tmp_path):
    """a tile of constant bytes drives single strings to thousands of bytes; a noisy tile fills the 4 094-entry table many times"""
    geom = mb.Geom(0, 512, 0, 256, 256, 512)
    a = np.zeros((256, 512), dtype=np.float32)
    a[:, 256:] = np.random.default_rng(1).standard_normal((256, 256)).astype(np.float32)
    p = str(tmp_path / "z.tif")
    mb.write_raster(p, geom, a, compression="LZW")
    assert os.path.getsize(p) < a.nbytes            # the zero tile compresses
    assert same(mb.read_raster(p)[1], a)
    assert same(np.array(PIL.open(p)), a)


@pytest.mark.parametrize("mode,dtype,comp", [("L", np.uint8, "raw"), ("I;16", np.uint16, "tiff_lzw"), ("I", np.int32, "packbits"),
                                             ("F", np.float32, "tiff_lzw"), ("I;16", np.uint16, "raw")])
def test_reads_striped_files_written_by_libtiff(tmp_path, mode, dtype, comp):
    rng = np.random.default_rng(3)
    a = (rng.integers(0, 200, (123, 211)) if dtype != np.float32 else rng.standard_normal((123, 211))).astype(dtype)
    a[10:30] = a[10, 0]
    p = str(tmp_path / "s.tif")
    PIL.fromarray(a).save(p, compression=comp)
    info = mb.raster_info(p)
    assert not info.tiled and not info.georeferenced and info.geom.as_tuple() == (0.0, 211.0, 0.0, 123.0, 123, 211)
    assert np.array_equal(mb.read_raster(p)[1], a.astype(np.float32))


def test_horizontal_predictor_and_band_selection(tmp_path):
    rng = np.random.default_rng(4)
    rgb = rng.integers(0, 255, (64, 80, 3)).astype(np.uint8)
    rgb[:, 20:60] = rgb[:, 20:21]
    p = str(tmp_path / "rgb.tif")
    PIL.fromarray(rgb).save(p, compression="tiff_lzw", tiffinfo={317: 2})
    info = mb.raster_info(p)
    assert info.nbands == 3 and info.predictor == 2
    for b in range(3):
        assert np.array_equal(mb.read_raster(p, band=b)[1], rgb[:, :, b].astype(np.float32))
    with pytest.raises(MbError):
        mb.read_raster(p, band=3)
    a16 = rng.integers(0, 60000, (50, 70)).astype(np.uint16)
    p2 = str(tmp_path / "p2.tif")
    PIL.fromarray(a16).save(p2, compression="tiff_lzw", tiffinfo={317: 2})
    assert mb.raster_info(p2).predictor == 2
    assert np.array_equal(mb.read_raster(p2)[1], a16.astype(np.float32))


def test_big_endian_strip_file(tmp_path):
    """hand-assembled MM file: 3 x 2 INT16, one strip, NoData 7"""
    vals = [1, -2, 7, 300, -32768, 5]
    data = struct.pack(">6h", *vals)
    nd = b"7\0"
    ents = [(256, 3, 1, 3 << 16), (257, 3, 1, 2 << 16), (258, 3, 1, 16 << 16), (259, 3, 1, 1 << 16), (262, 3, 1, 1 << 16),
            (273, 4, 1, 8), (277, 3, 1, 1 << 16), (278, 3, 1, 2 << 16), (279, 4, 1, 12), (339, 3, 1, 2 << 16),
            (42113, 2, 2, struct.unpack(">I", nd + b"\0\0")[0])]
    ifd = struct.pack(">H", len(ents)) + b"".join(struct.pack(">HHII", *e) for e in ents) + struct.pack(">I", 0)
    p = str(tmp_path / "be.tif")
    open(p, "wb").write(b"MM" + struct.pack(">HI", 42, 8 + len(data)) + data + ifd)
    a = mb.read_raster(p)[1]
    ref = np.array(vals, dtype=np.float32).reshape(2, 3)
    ref[ref == 7] = np.nan
    assert same(a, ref)


def test_refusals_are_errors_with_messages(tmp_path):
    with pytest.raises(MbError, match="cannot open"):
        mb.raster_info(str(tmp_path / "missing.tif"))
    p = str(tmp_path / "junk.tif")
    open(p, "wb").write(b"not a tiff file at all")
    with pytest.raises(MbError, match="not a TIFF"):
        mb.raster_info(p)
    a = np.arange(64 * 64, dtype=np.uint8).reshape(64, 64)
    p2 = str(tmp_path / "deflate.tif")
    PIL.fromarray(a).save(p2, compression="tiff_adobe_deflate")
    with pytest.raises(MbError, match="compression scheme"):
        mb.read_raster(p2)
    open(p, "wb").write(b"II" + struct.pack("<HQ", 43, 16) + b"\0" * 16)
    with pytest.raises(MbError, match="BigTIFF"):
        mb.raster_info(p)
    with pytest.raises(ValueError):
        mb.write_raster(p, mb.Geom(0, 1, 0, 1, 4, 4), np.zeros((3, 4), np.float32))


def test_read_stack_checks_the_grids(tmp_path):
    g = mb.Geom(10.0, 10.5, -2.0, -1.8, 20, 50)
    a, b = np.full((20, 50), 1.5, np.float32), np.full((20, 50), -4.0, np.float32)
    mb.write_raster(str(tmp_path / "a.tif"), g, a)
    mb.write_raster(str(tmp_path / "b.tif"), g, b)
    geom, planes = mb.read_stack([str(tmp_path / "a.tif"), str(tmp_path / "b.tif")])
    assert planes.shape == (2, 20, 50) and planes.dtype == np.float32 and same(planes[0], a) and same(planes[1], b)
    assert np.allclose(geom.as_tuple()[:4], g.as_tuple()[:4])
    mb.write_raster(str(tmp_path / "c.tif"), mb.Geom(10.0, 10.6, -2.0, -1.8, 20, 50), a)
    with pytest.raises(ValueError):
        mb.read_stack([str(tmp_path / "a.tif"), str(tmp_path / "c.tif")])


def test_write_geotiff_names_layers_like_the_reference(tmp_path):
    """machisplin.write.geotiff (V73:998-1022): <layer>.tif per response, or out.names"""
    g = mb.Geom(0.0, 4.0, 0.0, 3.0, 3, 4)
    layers = {"bio1": np.arange(12, dtype=np.float64).reshape(3, 4), "bio12": np.full((3, 4), np.nan)}
    paths = mb.write_geotiff(layers, g, out_dir=str(tmp_path))
    assert [os.path.basename(p) for p in paths] == ["bio1.tif", "bio12.tif"]
    assert same(mb.read_raster(paths[0])[1], layers["bio1"].astype(np.float32))
    assert np.isnan(mb.read_raster(paths[1])[1]).all()
    paths = mb.write_geotiff(layers, g, out_names=["temp", "prec"], out_dir=str(tmp_path))
    assert [os.path.basename(p) for p in paths] == ["temp.tif", "prec.tif"]
    with pytest.raises(ValueError):
        mb.write_geotiff(layers, g, out_names=["only_one"], out_dir=str(tmp_path))


def test_corrupted_files_fail_cleanly(tmp_path):
    """random byte flips and truncations of the golden fixture: an error or a raster, never a crash or an out-of-bounds write"""
    src = open(os.path.join(GOLD, "twi_2x2_tiles.tif"), "rb").read()
    ifd_at = int.from_bytes(src[4:8], "little")
    rng = np.random.default_rng(11)
    p = str(tmp_path / "f.tif")
    outcomes = {"ok": 0, "err": 0}
    for it in range(60):
        b = bytearray(src)
        mode = it % 4
        if mode == 0:
            for _ in range(rng.integers(1, 6)):
                b[rng.integers(ifd_at - 200, len(b))] = rng.integers(0, 256)
        elif mode == 1:
            for _ in range(rng.integers(1, 40)):
                b[rng.integers(8, ifd_at - 200)] = rng.integers(0, 256)
        elif mode == 2:
            b = b[:rng.integers(8, len(b))]
        else:
            b[rng.integers(0, 8)] = rng.integers(0, 256)
        open(p, "wb").write(bytes(b))
        try:
            info = mb.raster_info(p)
            if info.geom.nrow * info.geom.ncol <= 1_000_000:
                mb.read_raster(p)
            outcomes["ok"] += 1
        except (MbError, MemoryError):
            outcomes["err"] += 1
    assert outcomes["err"] > 0 and outcomes["ok"] + outcomes["err"] == 60
