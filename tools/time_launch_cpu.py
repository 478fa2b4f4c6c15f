#!/usr/bin/env python3
"""Host-side cost of one dctc_energy_batch_dev call (Python ctypes caller): wall clock of issuing launches of a tiny frame
without synchronising, next to the device time per launch."""
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dct_carver_b200 as dc  # noqa: E402
ctx = dc.Context(0)
w, ch, hmax = 3840, 3, 64
d_in = ctx.dev_alloc(w * hmax * ch)
d_out = ctx.dev_alloc(w * hmax * 4)
ctx.synth_fill_dev(d_in, 1, w * hmax * ch, w, hmax, ch, w * ch, 77, 0)
for b in (8, 4, 2, 16):
    ctx.set_params(b, 0.5, 0.5)
    for h in (8, 64):
        for _ in range(20):
            ctx.energy_batch_dev(d_in, 1, w * h * ch, w, h, ch, w * ch, d_out, w * h, w)
        ctx.sync()
        n = 2000
        ctx.timer_begin()
        t0 = time.perf_counter()
        for _ in range(n):
            ctx.energy_batch_dev(d_in, 1, w * h * ch, w, h, ch, w * ch, d_out, w * h, w)
        t1 = time.perf_counter()
        dev = ctx.timer_end() * 1e3 / n
        print("b=%2d h=%2d: host %.2f us per call, device %.2f us per launch" % (b, h, (t1 - t0) * 1e6 / n, dev))
ctx.close()
