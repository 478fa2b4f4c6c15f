#!/usr/bin/env python3
"""bench.py — energy-map throughput of the DCT-Carver hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload 4k|batch1080p|gigapixel]

A "step" is one pass of the hot path over one batch of synthetic frames that is larger than L2:
  4k          (default; BASELINE configs[1]) F distinct 3840x2160 RGB frames per GPU, one batched K1 launch
  batch1080p  (configs[3]) F distinct 1920x1080 RGB frames per GPU, frames sharded over ranks, no collective
  gigapixel   (configs[4]) one 32768x32768 RGB image split into row bands over the ranks; the halo rows are read
              straight from the neighbour rank's HBM over NVLink (CUDA IPC peer pointers) inside the K1 kernel
One process per GPU; for N>1 launch through torch.distributed.run (NCCL is used only for the barrier and the
max-over-ranks of the device time).  Rank 0 prints ONE JSON line.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "energy_map_throughput"
UNIT = "Mpix/s"
BYTES_PER_PX = {3: 7.0, 1: 5.0, 4: 8.0, 2: 6.0}   # u8 channels in + float32 energy out
SEED = 0xD0C7CA13

WORKLOADS = {
    "4k": dict(w=3840, h=2160, ch=3, frames=16, desc="3840x2160 RGB full energy map, blocksize 8, edges=textures=0.5"),
    "batch1080p": dict(w=1920, h=1080, ch=3, frames=64, desc="batch of 1920x1080 RGB frames, full energy maps"),
    "gigapixel": dict(w=32768, h=32768, ch=3, frames=1, desc="32768x32768 RGB, row bands + NVLink peer halo reads"),
}


def kernel_name(blocksize, kernel):
    if blocksize == 8:
        return {0: "dctc_k1_tc8_kernel (tcgen05 y-pass)", 3: "dctc_k1_tc8_kernel (tcgen05 y-pass)",
                2: "dctc_k1_march8_kernel (FP32x2 register march)"}.get(kernel, "dctc_k1_tile_kernel (FP32)")
    if blocksize in (2, 4) and kernel == 0:
        return "dctc_k1_small_kernel (FP32 streaming register march)"
    return "dctc_k1_tile_kernel (FP32)"


def arithmetic(blocksize, kernel):
    if blocksize == 8 and kernel in (0, 3):
        return ("exact integer luma, FP32 x-pass, y-pass on tcgen05 with fp16 hi/lo split operands (22 significant "
                "bits) and FP32 accumulation; max rel err vs the double reference 1.2e-6")
    return "FP32 (max rel err vs the double reference 1e-6)"


def kernel_note(blocksize):
    base = "traffic = DRAM bytes per launch from ncu (profiles/traffic.json); "
    if blocksize == 8:
        return base + ("blocksize 8 is compute-bound (24 tcgen05 MMAs + 32 FMNMX3 per 8x128 px), not HBM-bound: "
                       "see DESIGN.md section 4")
    if blocksize in (2, 4):
        return base + "blocksize %d: HBM / instruction-issue bound streaming kernel, see DESIGN.md section 4" % blocksize
    return base + "blocksize 16 runs in the FP32 tile kernel (FP32-pipe bound)"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.004)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def cpu_reference_rate(wl, seconds_target=12.0, frame=0):
    """Times the reference's own CPU path (oracle/_ref when it was compiled, else the oracle port) on a bounded
    sample of the workload: a block of full-width rows of one synthetic frame, all host threads."""
    import oracle_lib as ol
    cores = os.cpu_count() or 1
    kind = "reference" if ol.ref() is not None else "port"
    w, ch = wl["w"], wl["ch"]
    fn = ol.ref_energy if kind == "reference" else ol.oracle_energy

    def run(rows):
        hh = rows + 16
        img = ol.synth_image(w, hh, ch, SEED, 0, frame=frame, y_offset=1000)
        t0 = time.perf_counter()
        fn(img, 8, 0.5, 0.5, nthreads=cores)
        return (time.perf_counter() - t0), w * hh

    t, px = run(max(cores, 16))
    rate = px / t
    rows = int(min(max(rate * seconds_target / w, cores), 8192))
    t, px = run(rows)
    return dict(value=px / t / 1e6, unit=UNIT, cores=cores, kind=kind,
                sample="%d full-width rows (%d px) of one %dx%d %s frame, %d threads, %.1f s"
                       % (rows + 16, px, wl["w"], wl["h"], "RGB" if ch == 3 else "grey", cores, t)), t, px


def run_reference(args, rank, world):
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    vals = []
    total_t = 0.0
    target = max(1.0, min(12.0, 150.0 / max(1, args.steps + args.warmup)))
    for s in range(args.warmup + args.steps):
        cb, t, px = cpu_reference_rate(wl, seconds_target=target, frame=s)
        if s >= args.warmup:
            vals.append(px / t / 1e6)
            total_t += t
    v = statistics.mean(vals)
    cb["value"] = v
    line = {
        "metric": METRIC, "value": v, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total_t / max(1, len(vals)), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["desc"], "blocksize": 8, "edges": 0.5, "textures": 0.5,
                   "note": "reference CPU path (src/dct.c + fft2d via the LqrEnergyFunc callback), host threads only"},
        "cpu_baseline": cb,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="4k", choices=sorted(WORKLOADS))
    ap.add_argument("--frames", type=int, default=0, help="frames per GPU per step (default per workload)")
    ap.add_argument("--kernel", type=int, default=0)
    ap.add_argument("--blocksize", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    dist = None
    torch = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import dct_carver_b200 as dc
    wl = dict(WORKLOADS[args.workload])
    if args.frames:
        wl["frames"] = args.frames
    w, h, ch, F = wl["w"], wl["h"], wl["ch"], wl["frames"]
    ctx = dc.Context(local_rank, blocksize=args.blocksize, edges=0.5, textures=0.5, kernel=args.kernel)
    hbm_peak, peak_kind = peaks()

    def barrier():
        ctx.sync()
        if dist is not None:
            torch.cuda.synchronize()
            dist.barrier()

    if args.workload == "gigapixel":
        from dct_carver_b200 import multigpu
        runner = multigpu.BandRunner(ctx, dist, rank, world, w, h, ch, SEED)
        step = runner.step
        px_per_step_rank = runner.band_rows * w
        launches_per_step = 1
        cfg_extra = {"band_rows_per_gpu": runner.band_rows, "halo": "NVLink peer loads via CUDA IPC" if world > 1 else "none (single band)"}
    else:
        pitch = w * ch
        fstride = pitch * h
        d_in = ctx.dev_alloc(F * fstride)
        d_out = ctx.dev_alloc(F * w * h * 4)
        ctx.synth_fill_dev(d_in, F, fstride, w, h, ch, pitch, SEED, 0, first_frame=rank * F)
        ctx.sync()

        def step():
            ctx.energy_batch_dev(d_in, F, fstride, w, h, ch, pitch, d_out, w * h, w)
        px_per_step_rank = F * w * h
        launches_per_step = 1
        cfg_extra = {"frames_per_gpu_per_step": F}

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    step()      # one more untimed step after the barrier: the start event below is recorded behind it on the stream, so
                # the K timed steps run back to back on a busy GPU instead of starting from the idle state of the barrier
    l0 = ctx.launches
    ctx.timer_begin()
    for _ in range(args.steps):
        step()
    ms = ctx.timer_end()
    launches = ctx.launches - l0
    barrier()
    clocks = sampler.stop()
    if dist is not None:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        lt = torch.tensor([launches], device="cuda", dtype=torch.int64)
        dist.all_reduce(lt)
        launches = int(lt.item())
    ms_per_step = ms / args.steps
    total_px = px_per_step_rank * world
    value = total_px / (ms_per_step * 1e-3) / 1e6

    # roofline of the dominant kernel (K1): algorithmic bytes per launch / average launch time on one GPU
    alg_bytes = BYTES_PER_PX[ch] * px_per_step_rank
    achieved = alg_bytes / (ms_per_step * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            per_px = tj.get("%s_b%d_k%d" % (args.workload, args.blocksize, args.kernel), {}).get("dram_bytes_per_px")
            traffic = per_px * px_per_step_rank if per_px else None
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": traffic, "peak_kind": "of " + peak_kind, "algorithmic_bytes_per_launch": alg_bytes,
                "kernel": kernel_name(args.blocksize, args.kernel), "note": kernel_note(args.blocksize)}

    # e2e: the same metric through the host-buffer C-ABI call, pinned host memory, copies inside the timed region
    e2e = None
    if not args.no_e2e and args.workload != "gigapixel":
        Fe = min(F, 8)
        h_in = dc.pinned_array((Fe, h, w, ch), np.uint8)
        h_out = dc.pinned_array((Fe, h, w), np.float32)
        tmp = np.empty((h, w, ch), np.uint8)
        for f in range(Fe):
            ctx.d2h(tmp, d_in + f * fstride)
            h_in[f] = tmp
        for _ in range(2):
            ctx.energy_batch(h_in, h_out)
        barrier()
        ke = max(3, min(args.steps, 10))
        t0 = time.perf_counter()
        for _ in range(ke):
            ctx.energy_batch(h_in, h_out)
        te = (time.perf_counter() - t0) / ke
        if dist is not None:
            t = torch.tensor([te], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            te = float(t.item())
        e2e = {"value": Fe * w * h * world / te / 1e6, "unit": UNIT, "h2d_bytes_per_step": Fe * h * w * ch * world,
               "d2h_bytes_per_step": Fe * h * w * 4 * world, "frames_per_step": Fe * world,
               "api": "dctc_energy_batch (host buffers, 3-slot H2D/compute/D2H overlap)"}
        assert float(np.abs(h_out[0]).max()) > 0.0

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline, _, _ = cpu_reference_rate(wl)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong" if args.workload == "gigapixel" else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": dict({"workload": wl["desc"].replace("blocksize 8", "blocksize %d" % args.blocksize), "blocksize": args.blocksize, "edges": 0.5, "textures": 0.5,
                            "kernel": args.kernel, "arithmetic": arithmetic(args.blocksize, args.kernel),
                            "l2": "inputs+outputs per step exceed L2 (distinct frames)",
                            "parallelism": "frames sharded, no collective" if args.workload != "gigapixel" else "row bands"},
                           **cfg_extra),
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "clocks": clocks, "gpu_launches": launches,
        }
        print(json.dumps(line))
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
