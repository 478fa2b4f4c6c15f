#!/usr/bin/env python3
"""GPU check of the streaming kernel for block sizes 2 and 4 (dctc_k1_small.cu): bit-identity with the tile kernel and
parity with the oracle on several shapes / patterns, then device-timed 4K RGB throughput of both kernels."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import dct_carver_b200 as dc  # noqa: E402
import oracle_lib as ol  # noqa: E402


def main():
    ctx = dc.Context(0)
    bad = 0
    for b in (2, 4):
        for (pattern, ch, w, h) in [(0, 3, 128, 8), (0, 3, 200, 150), (0, 1, 131, 67), (3, 3, 96, 130), (1, 3, 640, 300),
                                    (2, 3, 515, 77), (0, 3, 1, 1), (0, 1, 3, 2), (0, 3, 7, 40), (0, 3, 1920, 1080), (0, 1, 1000, 700)]:
            for wts in [(0.5, 0.5), (0.8, 0.2)]:
                img = ol.synth_image(w, h, ch, 4321, pattern)
                ctx.set_params(b, *wts)
                ctx.set_kernel(dc.KERNEL_AUTO)
                got = ctx.energy_full(img)
                ctx.set_kernel(dc.KERNEL_FP32_TILE)
                tile = ctx.energy_full(img)
                same = bool((got.view(np.uint32) == tile.view(np.uint32)).all())
                msg = ""
                if w * h <= 700 * 400:
                    want = ol.oracle_energy(img, b, *wts).astype(np.float64)
                    err = np.abs(got.astype(np.float64) - want)
                    nbad = int((err > ol.ABS_TOL + ol.REL_TOL * np.abs(want)).sum())
                    msg = "max abs err vs oracle %.3e, out-of-tol %d" % (err.max(), nbad)
                    if wts[0] == wts[1] and nbad:
                        bad += 1
                print("b=%d pattern %d ch %d %4dx%-4d wts %s: bitwise == tile kernel: %s  %s" % (b, pattern, ch, w, h, wts, same, msg))
                if not same:
                    bad += 1
    # timing on 4K RGB, distinct frames
    F, n = 8, 10
    w, h, ch = 3840, 2160, 3
    d_in = ctx.dev_alloc(F * w * h * ch)
    d_out = ctx.dev_alloc(F * w * h * 4)
    ctx.synth_fill_dev(d_in, F, w * h * ch, w, h, ch, w * ch, 77, 0)
    for b in (2, 4):
        ctx.set_params(b, 0.5, 0.5)
        for name, k in (("tile  ", dc.KERNEL_FP32_TILE), ("stream", dc.KERNEL_AUTO)):
            ctx.set_kernel(k)
            for _ in range(2):
                ctx.energy_batch_dev(d_in, F, w * h * ch, w, h, ch, w * ch, d_out, w * h, w)
            ctx.sync()
            ctx.timer_begin()
            for _ in range(n):
                ctx.energy_batch_dev(d_in, F, w * h * ch, w, h, ch, w * ch, d_out, w * h, w)
            us = ctx.timer_end() * 1e3 / (n * F)
            print("b=%d %s: %.1f us per 4K frame, %.1f Gpix/s, %.1f%% of 6550.4 GB/s" % (b, name, us, w * h / us / 1e3, 100 * w * h * 7 / (us * 1e-6) / 6550.4e9))
    ctx.close()
    print("CHECK_SMALL", "PASS" if bad == 0 else "FAIL (%d)" % bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
