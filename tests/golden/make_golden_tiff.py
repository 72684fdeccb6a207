#!/usr/bin/env python
"""Builds tests/golden/twi_2x2_tiles.tif + twi_2x2_tiles.npz from the reference's bundled raster inst/extdata/TWI.tif
(INT16, 128 x 128 tiles, LZW as written by GDAL, GDAL_NODATA = -32768).

The fixture is a valid 256 x 256 TIFF whose four tiles are the ORIGINAL compressed bytes of tiles (0,0), (0,1), (1,0), (1,1) of
the reference file - LZW tiles are independent, so no re-encoding is involved - with the georeferencing tags of the NW corner.
Expected values: the same window of the full raster as decoded by Pillow / libtiff.  Run in the build container only
(/root/reference is not on the GPU box); the outputs are committed."""
import struct
import sys

import numpy as np
from PIL import Image

SRC = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/inst/extdata/TWI.tif"
OUT = __file__.rsplit("/", 1)[0]

im = Image.open(SRC)
t = im.tag_v2
W, H, TW, TH = t[256], t[257], t[322], t[323]
across = (W + TW - 1) // TW
offs, cnts = t[324], t[325]
raw = open(SRC, "rb").read()
tiles = [raw[offs[i]:offs[i] + cnts[i]] for i in (0, 1, across, across + 1)]
scale, tie, nodata = t[33550], t[33922], t[42113]

body = b""
toff, tcnt = [], []
pos = 8
for b in tiles:
    toff.append(pos)
    tcnt.append(len(b))
    body += b + (b"\0" if len(b) & 1 else b"")
    pos = 8 + len(body)
extra = b""


def put(data):
    global extra
    at = 8 + len(body) + len(extra)
    extra += data + (b"\0" if len(data) & 1 else b"")
    return at


o_off = put(struct.pack("<4I", *toff))
o_cnt = put(struct.pack("<4I", *tcnt))
o_scale = put(struct.pack("<3d", *scale))
o_tie = put(struct.pack("<6d", *tie))
nd = (str(nodata).strip() + "\0").encode()
o_nd = put(nd)
ents = [(256, 3, 1, 2 * TW), (257, 3, 1, 2 * TH), (258, 3, 1, 16), (259, 3, 1, 5), (262, 3, 1, 1), (277, 3, 1, 1),
        (284, 3, 1, 1), (317, 3, 1, 1), (322, 3, 1, TW), (323, 3, 1, TH), (324, 4, 4, o_off), (325, 4, 4, o_cnt),
        (339, 3, 1, 2), (33550, 12, 3, o_scale), (33922, 12, 6, o_tie), (42113, 2, len(nd), o_nd)]
ifd_at = 8 + len(body) + len(extra)
ifd = struct.pack("<H", len(ents)) + b"".join(struct.pack("<HHII", *e) for e in ents) + struct.pack("<I", 0)
open(f"{OUT}/twi_2x2_tiles.tif", "wb").write(b"II" + struct.pack("<HI", 42, ifd_at) + body + extra + ifd)

full = np.array(im)
np.savez_compressed(f"{OUT}/twi_2x2_tiles.npz", values=full[:2 * TH, :2 * TW].astype(np.int16), nodata=np.int32(int(float(nodata))),
                    xmin=tie[3], ymax=tie[4], rx=scale[0], ry=scale[1])
print("wrote", 8 + len(body) + len(extra) + len(ifd), "bytes")
