#!/usr/bin/env python3
"""Device-timed seam loop of BASELINE config 3 (1920x1080 RGB, 480 seams; the second 240 are timed with device events, as in
bench.py's configs.C3): python tools/time_seamloop.py   (DCTC_LIB=<path> loads an alternative build)"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import dct_carver_b200 as dc  # noqa: E402
if os.environ.get("DCTC_LIB"):
    dc.LIB_PATH = os.environ["DCTC_LIB"]
import oracle_lib as ol  # noqa: E402  (synthetic image generator only)
w, h, n = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1920, 1080, 480)
img = ol.synth_image(w, h, 3, 0xD0C7CA14, 0)
ctx = dc.Context(0, blocksize=8, edges=0.5, textures=0.5)
ctx.carver_load(img[:256, :256])
ctx.carver_resize_width(8)
ctx.carver_load(img)
first = ctx.carver_resize_width(n // 2)
ctx.timer_begin()
second = ctx.carver_resize_width(n - n // 2)
ms = ctx.timer_end()
print("%s %dx%d: %.1f us per seam (device time over the last %d seams), seam checksum %d" %
      (os.environ.get("DCTC_LIB", "libdctc.so")[-16:], w, h, 1e3 * ms / (n - n // 2), n - n // 2, int(np.concatenate([first, second]).astype(np.int64).sum())))
ctx.close()
