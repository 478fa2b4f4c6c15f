#!/usr/bin/env python3
"""K1-TC launch ramp: back-to-back single-frame launches of 3840-wide RGB frames of growing height (us per launch)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dct_carver_b200 as dc  # noqa: E402
if os.environ.get("DCTC_LIB"):
    dc.LIB_PATH = os.environ["DCTC_LIB"]
ctx = dc.Context(0)
w, ch = 3840, 3
b = int(sys.argv[1]) if len(sys.argv) > 1 else 8
F = 8
hmax = 2160
d_in = ctx.dev_alloc(F * w * hmax * ch)
d_out = ctx.dev_alloc(F * w * hmax * 4)
ctx.synth_fill_dev(d_in, F, w * hmax * ch, w, hmax, ch, w * ch, 77, 0)
ctx.set_params(b, 0.5, 0.5)
for h in (8, 16, 64, 240, 480, 1080, 2160):
    n = 40
    for i in range(4):
        ctx.energy_batch_dev(d_in + (i % F) * w * hmax * ch, 1, w * h * ch, w, h, ch, w * ch, d_out + (i % F) * w * hmax * 4, w * h, w)
    ctx.sync()
    ctx.timer_begin()
    for i in range(n):
        ctx.energy_batch_dev(d_in + (i % F) * w * hmax * ch, 1, w * h * ch, w, h, ch, w * ch, d_out + (i % F) * w * hmax * 4, w * h, w)
    us = ctx.timer_end() * 1e3 / n
    print("b=%d h=%4d: %.1f us per launch (%.1f Gpix/s)" % (b, h, us, w * h / us / 1e3))
ctx.close()
