#!/usr/bin/env python3
"""BASELINE config 3: 1920x1080 RGB retarget to 75 % width (480 seams) through the host carver with incremental
per-seam energy on the GPU (K2).  Reports us/seam and the split energy / cumulative map / seam search."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import dct_carver_b200 as dc  # noqa: E402
if os.environ.get("DCTC_LIB"):
    dc.LIB_PATH = os.environ["DCTC_LIB"]
from dct_carver_b200 import host  # noqa: E402
import oracle_lib as ol  # noqa: E402  (synthetic image generator only)


def main():
    w, h, n = 1920, 1080, 480
    if len(sys.argv) > 1:
        n = int(sys.argv[1])
    img = ol.synth_image(w, h, 3, 0xD0C7CA14, 0)
    ctx = dc.Context(0)
    host.render(img[:256, :256], -8, ctx=ctx)   # warm-up
    l0 = ctx.launches
    t0 = time.perf_counter()
    r = host.render(img, -n, 8, 0.5, 0.5, ctx=ctx, device_loop=False)
    dt = time.perf_counter() - t0
    lh = ctx.launches - l0
    # the same retarget with the whole seam loop on the device (dctc_carver_resize_width)
    ctx.set_params(8, 0.5, 0.5)
    ctx.carver_load(img[:256, :256])
    ctx.carver_resize_width(8)          # warm-up
    l1 = ctx.launches
    t1 = time.perf_counter()
    ctx.carver_load(img)
    t_load = time.perf_counter() - t1
    seams = ctx.carver_resize_width(n)
    dt_dev = time.perf_counter() - t1
    same = bool((seams == r["seams"]).all())
    # the drop-in path a reference user calls: dctc_render (mirror of render(), src/render.c:327-419) with the device loop
    t2 = time.perf_counter()
    r2 = host.render(img, -n, 8, 0.5, 0.5, ctx=ctx, device_loop=True)
    dt_render = time.perf_counter() - t2
    same = same and bool((r2["seams"] == r["seams"]).all()) and bool((r2["image"] == r["image"]).all())
    print(json.dumps({
        "dctc_render_device_loop_total_s": dt_render,
        "workload": "1920x1080 RGB -> %dx1080, %d vertical seams, blocksize 8" % (w - n, n),
        "device_loop_total_s": dt_dev, "device_loop_us_per_seam": 1e6 * (dt_dev - t_load) / n,
        "device_loop_load_and_full_map_s": t_load, "device_loop_gpu_launches": ctx.launches - l1,
        "device_loop_seams_equal_host_loop": same,
        "total_s": dt, "us_per_seam": 1e6 * dt / n, "seams_per_s": n / dt,
        "energy_s (K1 full + K2 band incl. H2D/D2H)": r["t_energy"], "cumulative_map_s (host)": r["t_mmap"],
        "seam_search_carve_s (host)": r["t_seam"], "gpu_launches": lh,
        "us_per_seam_energy": 1e6 * r["t_energy"] / n,
    }))


if __name__ == "__main__":
    main()
