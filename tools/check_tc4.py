#!/usr/bin/env python3
"""Quick GPU check of the block size 4 tensor-core kernel (K1-TC4, kernel id 3 with blocksize 4) against the oracle and the FP32 streaming kernel:
parity numbers for several shapes/patterns and a device-timed 4K throughput comparison."""
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
    for (pattern, ch, w, h) in [(0, 3, 128, 8), (0, 3, 200, 150), (0, 1, 131, 67), (3, 3, 96, 130), (1, 3, 640, 300),
                                (2, 3, 515, 77), (0, 3, 1, 1), (0, 1, 3, 2), (0, 3, 7, 40), (0, 3, 1920, 1080), (0, 3, 256, 16), (0, 1, 384, 24)]:
        for wts in [(0.5, 0.5), (0.8, 0.2)]:
            img = ol.synth_image(w, h, ch, 1234, pattern)
            ctx.set_params(4, *wts)
            ctx.set_kernel(dc.KERNEL_TC_SPLIT)
            try:
                got = ctx.energy_full(img)
            except dc.DctcError:
                print("CUDA error code", dc.lib().dctc_last_cuda_error(ctx.handle))
                raise
            ctx.set_kernel(dc.KERNEL_FP32_STREAM)
            ref32 = ctx.energy_full(img)
            want = ol.oracle_energy(img, 4, *wts) if w * h <= 700 * 400 else ref32
            g, r = got.astype(np.float64), want.astype(np.float64)
            err = np.abs(g - r)
            tol = ol.ABS_TOL + ol.REL_TOL * np.abs(r)
            nbad = int((err > tol).sum())
            e32 = np.abs(ref32.astype(np.float64) - r)
            print("pattern %d ch %d %4dx%-4d wts %s: max abs err %.3e (fp32 strm  %.3e) max rel %.3e  out-of-tol %d / %d"
                  % (pattern, ch, w, h, wts, err.max(), e32.max(), (err / np.maximum(np.abs(r), 1e-6)).max(), nbad, got.size))
            if wts[0] == wts[1]:
                bad += nbad
            if nbad and wts[0] == wts[1]:
                ys, xs = np.nonzero(err > tol)
                print("   first bad px:", list(zip(ys[:6].tolist(), xs[:6].tolist())), g[ys[0], xs[0]], r[ys[0], xs[0]])
    # timing, 4K RGB, 8 distinct frames per launch
    w, h, ch, F = 3840, 2160, 3, 8
    d_in = ctx.dev_alloc(F * w * h * ch)
    d_out = ctx.dev_alloc(F * w * h * 4)
    ctx.synth_fill_dev(d_in, F, w * h * ch, w, h, ch, w * ch, 77, 0)
    ctx.set_params(4, 0.5, 0.5)
    for k, name in [(dc.KERNEL_FP32_STREAM, "fp32 strm "), (dc.KERNEL_TC_SPLIT, "tcgen05   ")]:
        ctx.set_kernel(k)
        for _ in range(3):
            ctx.energy_batch_dev(d_in, F, w * h * ch, w, h, ch, w * ch, d_out, w * h, w)
        ctx.sync()
        ctx.timer_begin()
        n = 10
        for _ in range(n):
            ctx.energy_batch_dev(d_in, F, w * h * ch, w, h, ch, w * ch, d_out, w * h, w)
        ms = ctx.timer_end()
        us = ms * 1e3 / (n * F)
        print("%s: %.1f us per 4K frame, %.1f Gpix/s, %.1f%% of 6550 GB/s" % (name, us, w * h / us / 1e3, 100 * w * h * 7 / (us * 1e-6) / 6550.4e9))
    ctx.close()
    print("CHECK_TC4", "PASS" if bad == 0 else "FAIL")
    return 0 if bad == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
