"""Generates tests/golden/bundled_c1.npz from the only fixtures the reference ships (inputs only):
data-raw/sampling.csv (813 points) and the inst/extdata rasters' geometry + values at the points.
Run in the build container (needs /root/reference); the GPU box only reads the committed .npz.

The file pins (a) the inputs, (b) quantities derived by the oracle (regression pins - the reference
ships no expected outputs, parity vs R stays unpinned) and (c) numbers the survey derived
independently from V73:651-673 (2 x 3 internal tiles with 190/220/199/200/237/204 knots).
"""
import os
import sys

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import tiles as otl, tps as otps  # noqa: E402

REF = "/root/reference"
Image.MAX_IMAGE_PIXELS = None


def main():
    pts = np.genfromtxt(os.path.join(REF, "data-raw", "sampling.csv"), delimiter=",", skip_header=1)
    im = Image.open(os.path.join(REF, "inst", "extdata", "alt.tif"))
    ncol, nrow = im.size
    sx, sy, _ = im.tag_v2[33550]
    tie = im.tag_v2[33922]
    xmin, ymax = tie[3], tie[4]
    geom = (xmin, xmin + ncol * sx, ymax - nrow * sy, ymax, nrow, ncol)
    knots_xy, krow, kcol = otl.knot_coordinates(geom, pts[:, :2])
    planes = {}
    for name in ("alt", "slope", "TWI"):
        a = np.array(Image.open(os.path.join(REF, "inst", "extdata", name + ".tif")))
        planes[name] = a[krow, kcol].astype(np.float64)
    lay = otl.mltps_tile_layout(geom)
    counts = []
    fits = []
    y = pts[:, 2]                                     # bio_1 as the fitted field
    for fw in lay.fit_win:
        ins = (krow >= fw[0]) & (krow < fw[1]) & (kcol >= fw[2]) & (kcol < fw[3])
        counts.append(int(ins.sum()))
        f = otps.tps_fit(knots_xy[ins], y[ins])
        probe = otps.tps_predict_points(f, knots_xy[ins][:5] + 0.01)
        fits.append([f.lam, f.eff_df, f.d[0], f.d[1], f.d[2], np.abs(f.c).sum(), *probe])
    full = otps.tps_fit(knots_xy, y)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "bundled_c1.npz"),
                        points=pts, geom=np.array(geom), knots_xy=knots_xy, krow=krow, kcol=kcol,
                        alt=planes["alt"], slope=planes["slope"], twi=planes["TWI"],
                        fit_win=np.array(lay.fit_win), keep_win=np.array(lay.keep_win), tile_counts=np.array(counts),
                        tile_fits=np.array(fits),
                        full_fit=np.array([full.lam, full.eff_df, *full.d, np.abs(full.c).sum()]),
                        full_c=full.c)
    print("tiles", lay.nRx, lay.nCx, "knots per fit box", counts)
    print("full fit lambda", full.lam, "edf", full.eff_df, "sum|c|", np.abs(full.c).sum())


if __name__ == "__main__":
    main()
