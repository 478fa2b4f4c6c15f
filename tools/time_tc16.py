#!/usr/bin/env python3
"""Device-timed 4K throughput of the block-size-16 kernels (select an experimental build with DCTC_LIB=...)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dct_carver_b200 as dc  # noqa: E402


def main():
    ctx = dc.Context(0)
    w, h, ch = 3840, 2160, 3
    for F in (1, 16):
        d_in = ctx.dev_alloc(F * w * h * ch)
        d_out = ctx.dev_alloc(F * w * h * 4)
        ctx.synth_fill_dev(d_in, F, w * h * ch, w, h, ch, w * ch, 77, 0)
        ctx.set_params(16, 0.5, 0.5)
        for _ in range(2):
            ctx.energy_batch_dev(d_in, F, w * h * ch, w, h, ch, w * ch, d_out, w * h, w)
        ctx.sync()
        ctx.timer_begin()
        n = 5
        for _ in range(n):
            ctx.energy_batch_dev(d_in, F, w * h * ch, w, h, ch, w * ch, d_out, w * h, w)
        ms = ctx.timer_end()
        us = ms * 1e3 / (n * F)
        print("%s: %d frame(s)/launch: %.1f us per 4K frame, %.1f Gpix/s" % (os.environ.get("DCTC_LIB", "libdctc.so")[-14:], F, us, w * h / us / 1e3), flush=True)
        ctx.dev_free(d_in)
        ctx.dev_free(d_out)
    ctx.close()


if __name__ == "__main__":
    main()
