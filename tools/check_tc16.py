#!/usr/bin/env python3
"""Quick GPU check of the block-size-16 tensor-core kernel (dctc_k1_tc16.cu) against the compiled reference / oracle
and the FP32 tile kernel: parity numbers for several shapes, row bands with band_y0 (bit-equal to the full map), and
a device-timed 4K throughput comparison."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import dct_carver_b200 as dc  # noqa: E402
import oracle_lib as ol  # noqa: E402


def bands(ctx, img, bounds, b):
    h, w, ch = img.shape
    pitch = (w * ch + 15) & ~15
    buf = np.zeros((h, pitch), np.uint8)
    buf[:, :w * ch] = img.reshape(h, w * ch)
    d_img = ctx.dev_alloc(h * pitch)
    d_out = ctx.dev_alloc(h * w * 4)
    ctx.h2d(d_img, buf)
    rt, rb = b // 2 - 1, b // 2
    for (y0, y1) in bounds:
        t, bt = min(rt, y0), min(rb, h - y1)
        ctx.energy_band_dev(d_img + y0 * pitch, w, y1 - y0, ch, pitch, d_img + (y0 - t) * pitch if t else None, t, pitch,
                            d_img + y1 * pitch if bt else None, bt, pitch, d_out + y0 * w * 4, w, band_y0=y0)
    out = np.empty((h, w), np.float32)
    ctx.d2h(out, d_out)
    ctx.dev_free(d_img)
    ctx.dev_free(d_out)
    return out


def main():
    ctx = dc.Context(0)
    bad = 0
    quick = "--quick" in sys.argv
    shapes = [(0, 3, 64, 16), (0, 3, 128, 48), (0, 3, 200, 150), (0, 1, 131, 67), (3, 3, 96, 130), (1, 3, 640, 300), (2, 3, 515, 77),
              (0, 3, 1, 1), (0, 1, 3, 2), (0, 3, 7, 40), (0, 1, 1008, 40)]
    if not quick:
        shapes.append((0, 3, 1920, 1080))
    for (pattern, ch, w, h) in shapes:
        img = ol.synth_image(w, h, ch, 1234, pattern)
        ctx.set_params(16, 0.5, 0.5)
        ctx.set_kernel(dc.KERNEL_AUTO)
        n0 = ctx.launches
        try:
            got = ctx.energy_full(img)
        except dc.DctcError:
            print("CUDA error code", dc.lib().dctc_last_cuda_error(ctx.handle))
            raise
        ctx.set_kernel(dc.KERNEL_FP32_TILE)
        ref32 = ctx.energy_full(img)
        ctx.set_kernel(dc.KERNEL_AUTO)
        want = ol.best_energy(img, 16, 0.5, 0.5) if w * h <= 700 * 400 else ref32
        g, r = got.astype(np.float64), want.astype(np.float64)
        err = np.abs(g - r)
        tol = ol.ABS_TOL + ol.REL_TOL * np.abs(r)
        nbad = int((err > tol).sum())
        e32 = np.abs(ref32.astype(np.float64) - r)
        print("pattern %d ch %d %4dx%-4d: max abs err %.3e (fp32 tile %.3e) max rel %.3e  max |E| %.3f out-of-tol %d / %d"
              % (pattern, ch, w, h, err.max(), e32.max(), (err / np.maximum(np.abs(r), 1e-6)).max(), np.abs(r).max(), nbad, got.size), flush=True)
        bad += nbad
        if nbad:
            ys, xs = np.nonzero(err > tol)
            print("   first bad px:", list(zip(ys[:6].tolist(), xs[:6].tolist())), g[ys[0], xs[0]], r[ys[0], xs[0]])
        if h >= 40:
            cut = [(0, 13), (13, 29), (29, 30), (30, h)]
            gb = bands(ctx, img, cut, 16)
            same = np.array_equal(gb, got)
            print("   bands %s with band_y0: %s" % (cut, "bit-equal" if same else "DIFFER (max %.3e)" % np.abs(gb - got).max()), flush=True)
            bad += 0 if same else 1
    # timing, 4K RGB
    w, h, ch = 3840, 2160, 3
    for F in (1, 8):
        d_in = ctx.dev_alloc(F * w * h * ch)
        d_out = ctx.dev_alloc(F * w * h * 4)
        ctx.synth_fill_dev(d_in, F, w * h * ch, w, h, ch, w * ch, 77, 0)
        ctx.set_params(16, 0.5, 0.5)
        for k, name in [(dc.KERNEL_FP32_TILE, "fp32 tile "), (dc.KERNEL_AUTO, "tcgen05   ")]:
            ctx.set_kernel(k)
            for _ in range(2):
                ctx.energy_batch_dev(d_in, F, w * h * ch, w, h, ch, w * ch, d_out, w * h, w)
            ctx.sync()
            ctx.timer_begin()
            n = 5
            for _ in range(n):
                ctx.energy_batch_dev(d_in, F, w * h * ch, w, h, ch, w * ch, d_out, w * h, w)
            ms = ctx.timer_end()
            us = ms * 1e3 / (n * F)
            print("%s %d frame(s)/launch: %.1f us per 4K frame, %.1f Gpix/s, %.2f%% of 6550 GB/s"
                  % (name, F, us, w * h / us / 1e3, 100 * w * h * 7 / (us * 1e-6) / 6550.4e9), flush=True)
        ctx.set_kernel(dc.KERNEL_AUTO)
        ctx.dev_free(d_in)
        ctx.dev_free(d_out)
    ctx.close()
    print("CHECK_TC16", "PASS" if bad == 0 else "FAIL")
    return 0 if bad == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
