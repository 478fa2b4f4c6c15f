#!/usr/bin/env python3
"""Statistics of the incremental cumulative-map walk (needs a -DDCTC_INCR_STATS build of libdctc.so, path in DCTC_LIB):
average recomputed range width and walk cycles per row for a few synthetic patterns."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import dct_carver_b200 as dc
if os.environ.get("DCTC_LIB"):
    dc.LIB_PATH = os.environ["DCTC_LIB"]
import oracle_lib as ol
for pattern in (0, 3):
    img = ol.synth_image(1920, 1080, 3, 0xD0C7CA14, pattern)
    ctx = dc.Context(0); ctx.set_params(8, 0.5, 0.5); ctx.carver_load(img); ctx.carver_set_incremental(True)
    ctx.carver_resize_width(100)
    print("pattern", pattern, "rebuilds", ctx.carver_rebuild_count())
    ctx.close()
