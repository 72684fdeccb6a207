#!/usr/bin/env python
"""GeoTIFF -> float32 plane in HBM: host decoder + upload against the device decoder (mb_tiff_read_f32_dev) on an 8192 x 8192 raster
(FLT4S as terra::writeRaster writes it, LZW and uncompressed; INT16 LZW like the reference's covariates).

    python tools/tiff_check.py [--n 8192]
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                                                     # noqa: E402  (device buffers only)
import machisplin_b200 as mb                                     # noqa: E402
from machisplin_b200 import geotiff, synth                       # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=8192)
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()
eng = mb.Engine(0)
geom = synth.make_geom(args.n, args.n)
plane = synth.covariate_planes(geom, 1)[0]
tmp = tempfile.mkdtemp()
dev = torch.empty((geom.nrow, geom.ncol), dtype=torch.float32, device="cuda:0")
pinned = torch.empty((geom.nrow, geom.ncol), dtype=torch.float32, pin_memory=True)
cases = []
for comp in ("NONE", "LZW"):
    p = os.path.join(tmp, f"f32_{comp}.tif")
    mb.write_raster(p, geom, plane, compression=comp)
    cases.append((f"FLT4S {comp}", p))
try:
    from PIL import Image
    Image.MAX_IMAGE_PIXELS = None
    q = np.nan_to_num(plane, nan=-32768.0).round().astype(np.int16)
    p = os.path.join(tmp, "i16_lzw.tif")
    Image.fromarray(q).save(p, compression="tiff_lzw")
    cases.append(("INT2S LZW strips (libtiff)", p))
except Exception as ex:                                           # Pillow is optional
    print("no INT16 case:", ex)
for name, p in cases:
    size = os.path.getsize(p)
    host = []
    for _ in range(args.reps):
        t0 = time.perf_counter()
        mb.read_raster(p, out=pinned.numpy())
        dev.copy_(pinned, non_blocking=True)
        torch.cuda.synchronize()
        host.append(time.perf_counter() - t0)
    ref = dev.clone()
    gpu = []
    for _ in range(args.reps):
        t0 = time.perf_counter()
        st = geotiff.read_raster_dev(eng, p, dev.data_ptr())
        torch.cuda.synchronize()
        gpu.append(time.perf_counter() - t0)
    same = bool(torch.equal(torch.nan_to_num(dev, nan=-7.0), torch.nan_to_num(ref, nan=-7.0)))
    print(json.dumps({"case": name, "file_MB": round(size / 1e6, 1), "plane_MB": round(plane.nbytes / 1e6, 1),
                      "host_decode_plus_upload_ms": round(1e3 * min(host), 1), "device_decode_ms": round(1e3 * min(gpu), 1),
                      "h2d_MB": round(st["h2d_bytes"] / 1e6, 1), "decoded_on_gpu": st["decoded_on_gpu"], "chunks": st["chunks"],
                      "identical": same, "host_threads": os.cpu_count()}), flush=True)
