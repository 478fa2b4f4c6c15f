#!/usr/bin/env python3
"""Multi-GPU correctness check, run under torchrun on N GPUs:
   python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/check_bands_multi.py
Each rank computes its row band of a synthetic image with halos read from the neighbours' HBM (peer mode) and by
explicit NCCL halo exchange; rank 0 recomputes the whole image alone and compares bit for bit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dct_carver_b200 as dc  # noqa: E402
from dct_carver_b200 import multigpu  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w, h, ch, seed = 4096, 2051, 3, 77
    ok = True
    for b in (8, 16, 4, 2):
        ctx = dc.Context(local, blocksize=b)
        for mode in ("peer", "exchange"):
            r = multigpu.BandRunner(ctx, dist, rank, world, w, h, ch, seed, mode=mode)
            r.step()
            ctx.sync()
            torch.cuda.synchronize()
            dist.barrier()
            band = r.fetch()
            t = torch.from_numpy(band).cuda()
            sizes = [y1 - y0 for y0, y1 in multigpu.band_bounds(h, world)]
            outs = [torch.empty((s, w), dtype=torch.float32, device="cuda") for s in sizes]
            dist.all_gather(outs, t) if len(set(sizes)) == 1 else [dist.broadcast(outs[i] if i != rank else t, i) for i in range(world)]
            if len(set(sizes)) != 1:
                outs[rank] = t
            if rank == 0:
                full = torch.cat(outs, 0).cpu().numpy()
                d_img = ctx.dev_alloc(w * h * ch)
                d_out = ctx.dev_alloc(w * h * 4)
                ctx.synth_fill_dev(d_img, 1, 0, w, h, ch, w * ch, seed, 0)
                ctx.energy_batch_dev(d_img, 1, 0, w, h, ch, w * ch, d_out, 0, w, sync=True)
                want = np.empty((h, w), np.float32)
                ctx.d2h(want, d_out)
                same = np.array_equal(full, want)
                ok &= same
                print("b=%d mode=%s world=%d bands==single-GPU: %s" % (b, mode, world, same), flush=True)
                ctx.dev_free(d_img)
                ctx.dev_free(d_out)
            dist.barrier()
            r.close()
        ctx.close()
    if rank == 0:
        print("MULTIGPU_CHECK", "PASS" if ok else "FAIL", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
