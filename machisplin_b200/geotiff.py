"""GeoTIFF in / out through the C ABI (``mb_tiff_*``, csrc/tiff_io.cu): the host-side mirror of ``terra::rast(path)``
(README Example 1, V73:68-70) and ``terra::writeRaster(x, filename)`` (V73:1011, 1020).  No GPU is needed for these calls,
except ``read_raster_dev`` (decode on the device, plane born in HBM)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _lib
from ._lib import check
from .engine import Geom, as_geom


class _Meta(C.Structure):
    _fields_ = [("grid", _lib.Grid),
                ("nbands", C.c_int32), ("bits", C.c_int32), ("sample_format", C.c_int32), ("compression", C.c_int32),
                ("predictor", C.c_int32), ("tiled", C.c_int32), ("chunk_w", C.c_int32), ("chunk_h", C.c_int32),
                ("has_georef", C.c_int32), ("has_nodata", C.c_int32), ("epsg", C.c_int32), ("nodata", C.c_double)]


@dataclass(frozen=True)
class RasterInfo:
    """What ``terra::rast(path)`` prints: geometry, bands, data type, NoData flag (plus the file layout)."""
    geom: Geom
    nbands: int
    dtype: str            # "INT2S", "INT2U", "FLT4S", ... (terra's datatype names)
    compression: int      # TIFF code: 1 none, 5 LZW, 32773 PackBits
    predictor: int
    tiled: bool
    chunk: tuple          # (width, height) of a tile, or (ncol, rows per strip)
    georeferenced: bool
    nodata: Optional[float]
    epsg: int


_TERRA = {(1, 8): "INT1U", (2, 8): "INT1S", (1, 16): "INT2U", (2, 16): "INT2S", (1, 32): "INT4U", (2, 32): "INT4S",
          (3, 32): "FLT4S", (3, 64): "FLT8S"}


def raster_info(path: str) -> RasterInfo:
    m = _Meta()
    check(_lib.load().mb_tiff_info(str(path).encode(), C.byref(m)))
    g = m.grid
    return RasterInfo(Geom(g.xmin, g.xmax, g.ymin, g.ymax, g.nrow, g.ncol), m.nbands,
                      _TERRA.get((m.sample_format, m.bits), f"fmt{m.sample_format}/{m.bits}"), m.compression, m.predictor,
                      bool(m.tiled), (m.chunk_w, m.chunk_h), bool(m.has_georef), m.nodata if m.has_nodata else None, m.epsg)


def read_raster(path: str, band: int = 0, out: Optional[np.ndarray] = None, threads: int = 0):
    """One band as float32 in terra cell order (row-major from the NW corner), NoData -> NaN.  ``out`` may be a
    preallocated (pinned) float32 array of shape (nrow, ncol), e.g. one plane of the covariate stack handed to
    ``Engine.mltps_predict``.  Returns (Geom, array)."""
    info = raster_info(path)
    shape = (info.geom.nrow, info.geom.ncol)
    if out is None:
        out = np.empty(shape, dtype=np.float32)
    if out.dtype != np.float32 or out.shape != shape or not out.flags.c_contiguous:
        raise ValueError(f"out must be a C-contiguous float32 array of shape {shape}")
    check(_lib.load().mb_tiff_read_f32(str(path).encode(), int(band), out.ctypes.data_as(_lib.PF), int(threads)))
    return info.geom, out


class _DevStats(C.Structure):
    _fields_ = [("decoded_on_gpu", C.c_int32), ("chunks", C.c_int32), ("h2d_bytes", C.c_int64)]


def read_raster_dev(engine, path: str, out_ptr: int, band: int = 0, threads: int = 0, stream: int = 0) -> dict:
    """The band decoded ON THE DEVICE into the float32 plane at ``out_ptr`` (nrow * ncol floats in HBM, e.g. one plane of the
    covariate stack ``Engine.mltps_predict_dev`` reads): compressed bytes cross PCIe, the GPU undoes LZW, predictor, sample
    type and NoData (``mb_tiff_read_f32_dev``).  Returns {"decoded_on_gpu", "chunks", "h2d_bytes"}."""
    st = _DevStats()
    check(engine.lib.mb_tiff_read_f32_dev(engine._h, str(path).encode(), int(band), C.c_void_p(int(out_ptr)), int(threads),
                                          C.c_void_p(int(stream)), C.byref(st)))
    return {"decoded_on_gpu": bool(st.decoded_on_gpu), "chunks": int(st.chunks), "h2d_bytes": int(st.h2d_bytes)}


def read_stack(paths, threads: int = 0):
    """``c(rast(a), rast(b), ...)``: the covariate planes [C][nrow][ncol] float32 of rasters on one grid."""
    geom0, planes = None, None
    for k, p in enumerate(paths):
        info = raster_info(p)
        if geom0 is None:
            geom0 = info.geom
            planes = np.empty((len(paths), geom0.nrow, geom0.ncol), dtype=np.float32)
        elif (info.geom.nrow, info.geom.ncol) != (geom0.nrow, geom0.ncol) or \
                max(abs(a - b) for a, b in zip(info.geom.as_tuple()[:4], geom0.as_tuple()[:4])) > 1e-6 * geom0.rx:
            raise ValueError(f"{p}: extent or dimensions differ from the first raster (terra: 'extents do not match')")
        read_raster(p, 0, planes[k], threads)
    return geom0, planes


def write_raster(path: str, geom, values: np.ndarray, compression: str = "LZW", epsg: int = 0, threads: int = 0):
    """``terra::writeRaster(x, filename, overwrite=TRUE)``: FLT4S GeoTIFF, NaN = NoData."""
    geom = as_geom(geom)
    comp = {"NONE": 1, "LZW": 5}[compression.upper()]
    a = np.ascontiguousarray(values)
    if a.shape != (geom.nrow, geom.ncol):
        raise ValueError("values must have shape (nrow, ncol)")
    lib = _lib.load()
    g = geom.c()
    if a.dtype == np.float64:
        check(lib.mb_tiff_write_f64(str(path).encode(), C.byref(g), a.ctypes.data_as(_lib.PD), comp, int(epsg), int(threads)))
    else:
        a = np.ascontiguousarray(a, dtype=np.float32)
        check(lib.mb_tiff_write_f32(str(path).encode(), C.byref(g), a.ctypes.data_as(_lib.PF), comp, int(epsg), int(threads)))


def write_geotiff(layers, geom, out_names=None, out_dir: str = ".", compression: str = "LZW", epsg: int = 0):
    """``machisplin.write.geotiff(mltps.in, out.names)`` (V73:998-1022), raster part: one ``<name>.tif`` per response layer in
    ``out_dir`` (the reference writes into ``getwd()``), named after the layer unless ``out_names`` is given.
    ``layers`` maps layer name -> final raster (nrow x ncol), in response order.  Returns the paths written."""
    names = list(layers)
    if out_names is not None:
        out_names = list(out_names)
        if len(out_names) < len(names):
            raise ValueError("out_names is shorter than the number of layers")
    import os
    paths = []
    for i, name in enumerate(names):
        p = os.path.join(out_dir, f"{out_names[i] if out_names is not None else name}.tif")
        write_raster(p, geom, layers[name], compression=compression, epsg=epsg)
        paths.append(p)
    return paths
