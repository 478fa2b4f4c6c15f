#!/usr/bin/env python3
"""Times the streaming kernel for block size 2 or 4 on 4K RGB frames: python tools/time_small.py <b> [frames] [iters]"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dct_carver_b200 as dc  # noqa: E402
if os.environ.get("DCTC_LIB"):
    dc.LIB_PATH = os.environ["DCTC_LIB"]
b = int(sys.argv[1]) if len(sys.argv) > 1 else 2
F = int(sys.argv[2]) if len(sys.argv) > 2 else 8
n = int(sys.argv[3]) if len(sys.argv) > 3 else 10
ctx = dc.Context(0)
w, h, ch = 3840, 2160, 3
d_in = ctx.dev_alloc(F * w * h * ch)
d_out = ctx.dev_alloc(F * w * h * 4)
ctx.synth_fill_dev(d_in, F, w * h * ch, w, h, ch, w * ch, 77, 0)
ctx.set_params(b, float(os.environ.get('DCTC_EDGES', 0.5)), float(os.environ.get('DCTC_TEXTURES', 0.5)))
for _ in range(2):
    ctx.energy_batch_dev(d_in, F, w * h * ch, w, h, ch, w * ch, d_out, w * h, w)
ctx.sync()
ctx.timer_begin()
for _ in range(n):
    ctx.energy_batch_dev(d_in, F, w * h * ch, w, h, ch, w * ch, d_out, w * h, w)
us = ctx.timer_end() * 1e3 / (n * F)
print("b=%d stream: %.1f us per 4K frame, %.1f Gpix/s, %.1f%% of 6550.4 GB/s" % (b, us, w * h / us / 1e3, 100 * w * h * 7 / (us * 1e-6) / 6550.4e9))
ctx.close()
