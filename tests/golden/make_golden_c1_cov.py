"""Generates tests/golden/bundled_c1_cov4.npz: the reference's three bundled covariate rasters
(inst/extdata/{alt,slope,TWI}.tif, README Example 1) decoded by the library's own GeoTIFF reader
(mb_tiff_read_f32: NoData -> NaN) and decimated 4 x (the centre cell of every 4 x 4 block), so that the
reference's one real workload - rough INT16 terrain fields, 813 points - fits a committed fixture.
int16 storage, -32768 = NA.  Run in the build container (needs /root/reference); the GPU box reads the .npz.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from machisplin_b200 import geotiff  # noqa: E402

REF = "/root/reference/inst/extdata"
F = 4


def main():
    out = {}
    geom = None
    for name in ("alt", "slope", "TWI"):
        g, a = geotiff.read_raster(os.path.join(REF, name + ".tif"))
        nr, nc = (g.nrow // F) * F, (g.ncol // F) * F
        a = a[:nr, :nc][F // 2::F, F // 2::F]
        assert np.all(np.isnan(a) | (a == np.round(a))) and np.nanmax(np.abs(a)) < 32767
        q = np.where(np.isnan(a), -32768, a).astype(np.int16)
        out[name.lower()] = q
        geom = (g.xmin, g.xmin + nc * g.rx, g.ymax - nr * g.ry, g.ymax, nr // F, nc // F)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "bundled_c1_cov4.npz"), geom=np.array(geom), **out)
    print(geom, {k: (v.shape, float((v == -32768).mean())) for k, v in out.items()})


if __name__ == "__main__":
    main()
