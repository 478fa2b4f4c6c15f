#!/bin/bash
cat > /tmp/t4.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import dct_carver_b200 as dc
ctx = dc.Context(0)
w, h, ch = 3840, 2160, 3
for F, n in ((16, 10), (64, 5)):
    d_in = ctx.dev_alloc(F * w * h * ch); d_out = ctx.dev_alloc(F * w * h * 4)
    ctx.synth_fill_dev(d_in, F, w * h * ch, w, h, ch, w * ch, 77, 0)
    ctx.set_params(4, 0.5, 0.5)
    for k, name in ((dc.KERNEL_FP32_STREAM, "stream"), (dc.KERNEL_TC_SPLIT, "tc4")):
        ctx.set_kernel(k)
        for _ in range(2): ctx.energy_batch_dev(d_in, F, w * h * ch, w, h, ch, w * ch, d_out, w * h, w)
        ctx.sync(); ctx.timer_begin()
        for _ in range(n): ctx.energy_batch_dev(d_in, F, w * h * ch, w, h, ch, w * ch, d_out, w * h, w)
        us = ctx.timer_end() * 1e3 / (n * F)
        print("seg %s F=%d %s: %.1f us per 4K frame (%.1f%%)" % (os.environ.get("DCTC_TC4_SEG", "256"), F, name, us, 100 * w * h * 7 / (us * 1e-6) / 6550.4e9))
    ctx.dev_free(d_in); ctx.dev_free(d_out)
PY
for seg in 64 128 256 512 1088 2160; do DCTC_TC4_SEG=$seg timeout 120 python /tmp/t4.py | grep tc4; done
timeout 120 python /tmp/t4.py | grep stream
