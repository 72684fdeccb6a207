"""GeoTIFF decoded on the device (mb_tiff_read_f32_dev): the plane in HBM must be bit-identical to what the
host decoder (itself checked against libtiff and the reference's own raster in test_geotiff_cpu.py) returns for the same file -
LZW tiles exactly as GDAL wrote them into the reference's TWI.tif, what the writer produces, striped libtiff files, predictor 2,
ragged edges, long LZW strings and table resets - and the formats outside the GPU path must fall back, not fail."""
import os
import struct

import numpy as np
import pytest

import machisplin_b200 as mb
from machisplin_b200 import geotiff
from machisplin_b200._lib import MbError

pytestmark = pytest.mark.gpu
PIL = pytest.importorskip("PIL.Image")
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def same(a, b):
    return np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(np.nan_to_num(a).view(np.uint32), np.nan_to_num(b).view(np.uint32))


def dev_read(engine, path, band=0):
    import torch
    info = mb.raster_info(path)
    out = torch.full((info.geom.nrow, info.geom.ncol), -12345.0, dtype=torch.float32, device="cuda:0")
    st = geotiff.read_raster_dev(engine, path, out.data_ptr(), band=band)
    return out.cpu().numpy(), st


def test_golden_tiles_of_the_reference_raster_on_the_device(engine):
    p = os.path.join(GOLD, "twi_2x2_tiles.tif")
    got, st = dev_read(engine, p)
    assert st["decoded_on_gpu"] and st["chunks"] == 4 and st["h2d_bytes"] < got.nbytes // 2   # INT16 + LZW: far fewer bytes than the float plane
    assert same(got, mb.read_raster(p)[1])
    z = np.load(os.path.join(GOLD, "twi_2x2_tiles.npz"))
    ref = z["values"].astype(np.float32)
    ref[z["values"] == z["nodata"]] = np.nan
    assert same(got, ref)


@pytest.mark.parametrize("shape", [(1, 1), (256, 256), (300, 517), (700, 3), (257, 1025)])
@pytest.mark.parametrize("compression", ["NONE", "LZW"])
def test_what_the_writer_writes(engine, tmp_path, shape, compression):
    rng = np.random.default_rng(shape[0] * 7 + shape[1])
    a = rng.standard_normal(shape).astype(np.float32)
    a[rng.random(shape) < 0.05] = np.nan
    if shape[0] > 100:
        a[40:90] = 3.25                      # long runs: deep LZW strings, copied by the whole warp
    geom = mb.Geom(-77.5, -77.5 + 0.001 * shape[1], -6.0, -6.0 + 0.002 * shape[0], shape[0], shape[1])
    p = str(tmp_path / "r.tif")
    mb.write_raster(p, geom, a, compression=compression, epsg=4326)
    got, st = dev_read(engine, p)
    assert st["decoded_on_gpu"]
    assert same(got, a)


def test_lzw_table_resets_and_long_strings_on_the_device(engine, tmp_path):
    geom = mb.Geom(0, 512, 0, 256, 256, 512)
    a = np.zeros((256, 512), dtype=np.float32)
    a[:, 256:] = np.random.default_rng(1).standard_normal((256, 256)).astype(np.float32)
    p = str(tmp_path / "z.tif")
    mb.write_raster(p, geom, a, compression="LZW")
    got, st = dev_read(engine, p)
    assert st["decoded_on_gpu"] and same(got, a)


@pytest.mark.parametrize("mode,dtype,comp,on_gpu", [("L", np.uint8, "raw", True), ("I;16", np.uint16, "tiff_lzw", True),
                                                    ("I", np.int32, "packbits", False), ("F", np.float32, "tiff_lzw", True),
                                                    ("I;16", np.uint16, "raw", True), ("I", np.int32, "tiff_lzw", True)])
def test_striped_files_written_by_libtiff(engine, tmp_path, mode, dtype, comp, on_gpu):
    rng = np.random.default_rng(3)
    a = (rng.integers(0, 200, (123, 211)) if dtype != np.float32 else rng.standard_normal((123, 211))).astype(dtype)
    a[10:30] = a[10, 0]
    p = str(tmp_path / "s.tif")
    PIL.fromarray(a).save(p, compression=comp)
    got, st = dev_read(engine, p)
    assert st["decoded_on_gpu"] == on_gpu                      # PackBits: host decoder + upload
    assert same(got, mb.read_raster(p)[1]) and np.array_equal(got, a.astype(np.float32))


def test_predictor_2_and_interleaved_bands(engine, tmp_path):
    rng = np.random.default_rng(4)
    a16 = rng.integers(0, 60000, (150, 333)).astype(np.uint16)
    a16[:, 100:300] = a16[:, 100:101]
    p2 = str(tmp_path / "p2.tif")
    PIL.fromarray(a16).save(p2, compression="tiff_lzw", tiffinfo={317: 2})
    assert mb.raster_info(p2).predictor == 2
    got, st = dev_read(engine, p2)
    assert st["decoded_on_gpu"] and np.array_equal(got, a16.astype(np.float32))
    rgb = rng.integers(0, 255, (64, 80, 3)).astype(np.uint8)
    p = str(tmp_path / "rgb.tif")
    PIL.fromarray(rgb).save(p, compression="tiff_lzw")        # predictor 1, pixel-interleaved: band selection on the device
    for b in range(3):
        got, st = dev_read(engine, p, band=b)
        assert st["decoded_on_gpu"] and np.array_equal(got, rgb[:, :, b].astype(np.float32))
    PIL.fromarray(rgb).save(p, compression="tiff_lzw", tiffinfo={317: 2})   # predictor 2 across interleaved samples: host path
    for b in range(3):
        got, st = dev_read(engine, p, band=b)
        assert not st["decoded_on_gpu"] and np.array_equal(got, rgb[:, :, b].astype(np.float32))


def test_big_endian_file_falls_back_to_the_host_decoder(engine, tmp_path):
    vals = [1, -2, 7, 300, -32768, 5]
    data = struct.pack(">6h", *vals)
    nd = b"7\0"
    ents = [(256, 3, 1, 3 << 16), (257, 3, 1, 2 << 16), (258, 3, 1, 16 << 16), (259, 3, 1, 1 << 16), (262, 3, 1, 1 << 16),
            (273, 4, 1, 8), (277, 3, 1, 1 << 16), (278, 3, 1, 2 << 16), (279, 4, 1, 12), (339, 3, 1, 2 << 16),
            (42113, 2, 2, struct.unpack(">I", nd + b"\0\0")[0])]
    ifd = struct.pack(">H", len(ents)) + b"".join(struct.pack(">HHII", *e) for e in ents) + struct.pack(">I", 0)
    p = str(tmp_path / "be.tif")
    open(p, "wb").write(b"MM" + struct.pack(">HI", 42, 8 + len(data)) + data + ifd)
    got, st = dev_read(engine, p)
    assert not st["decoded_on_gpu"]
    assert same(got, mb.read_raster(p)[1])


def test_corrupt_lzw_stream_is_an_error_not_a_crash(engine, tmp_path):
    geom = mb.Geom(0, 256, 0, 256, 256, 256)
    a = np.random.default_rng(9).standard_normal((256, 256)).astype(np.float32)
    p = str(tmp_path / "c.tif")
    mb.write_raster(p, geom, a, compression="LZW")
    raw = bytearray(open(p, "rb").read())
    # flip bytes in the middle of the tile data (the IFD of the writer sits at the end of the file)
    for k in range(2000, 2100):
        raw[k] ^= 0xFF
    open(p, "wb").write(bytes(raw))
    try:
        host = mb.read_raster(p)[1]
        host_err = None
    except MbError as ex:
        host, host_err = None, ex
    try:
        got, st = dev_read(engine, p)
        dev_err = None
    except MbError as ex:
        got, dev_err = None, ex
    assert (host_err is None) == (dev_err is None)            # both decode (garbage in, same garbage out) or both refuse
    if host_err is None:
        assert same(got, host)


def test_stack_of_device_planes_feeds_mltps_predict_dev(engine, tmp_path):
    """The use the device reader exists for: covariate planes decoded straight into the [C][nrow][ncol] stack in HBM that
    mb_mltps_predict_dev reads - same raster as through host planes."""
    import torch
    from machisplin_b200 import synth
    geom = synth.make_geom(192, 256)
    C = 3
    cov = synth.covariate_planes(geom, C)
    paths = []
    for k in range(C):
        p = str(tmp_path / f"cov{k}.tif")
        mb.write_raster(p, geom, cov[k], compression="LZW")
        paths.append(p)
    stack = torch.empty((C, geom.nrow, geom.ncol), dtype=torch.float32, device="cuda:0")
    for k, p in enumerate(paths):
        geotiff.read_raster_dev(engine, p, stack[k].data_ptr())
    assert same(stack.cpu().numpy().reshape(-1, geom.ncol), cov.reshape(-1, geom.ncol))
    xy, _, _ = synth.make_knots(geom, 300, 5)
    resid = synth.residual_field(xy, 5)
    models = synth.make_models(geom, C, 300, 5, kept="gnmv")
    kept, w, wt = synth.ensemble_weights("gnmv")
    ens = engine.ensemble_create(geom, models, kept, w, wt, C + 2)
    out = torch.empty((geom.nrow, geom.ncol), dtype=torch.float64, device="cuda:0")
    engine.mltps_predict_dev(geom, ens, stack.data_ptr(), C, xy, resid, out.data_ptr())
    torch.cuda.synchronize()
    ref, _ = engine.mltps_predict(geom, ens, cov, xy, resid)
    assert np.array_equal(np.isnan(out.cpu().numpy()), np.isnan(ref))
    np.testing.assert_array_equal(np.nan_to_num(out.cpu().numpy()), np.nan_to_num(ref))
