#!/usr/bin/env python3
"""Multi-GPU correctness check of the C host layer on real peers.

  one process, all devices (dctc_multi_*):          python tools/check_multi.py
  one process per GPU (dctc_band_runner_*):         python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 \
                                                        --master-port 29511 tools/check_multi.py
Both compare the row-band result (halo rows read from the neighbour GPU's HBM over NVLink) bit for bit with one GPU,
for every block size; the single-process mode also checks the frame round-robin and the sharded energy image."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dct_carver_b200 as dc  # noqa: E402


def single_process():
    n = dc.lib().dctc_device_count()
    ok = True
    w, h, ch, seed = 4096, 2051, 3, 77
    for b in (8, 16, 4, 2):
        one = dc.Context(0, blocksize=b)
        d_img = one.dev_alloc(w * h * ch)
        d_out = one.dev_alloc(w * h * 4)
        one.synth_fill_dev(d_img, 1, 0, w, h, ch, w * ch, seed, 0)
        one.energy_batch_dev(d_img, 1, 0, w, h, ch, w * ch, d_out, 0, w, sync=True)
        want = np.empty((h, w), np.float32)
        img = np.empty((h, w, ch), np.uint8)
        one.d2h(want, d_out)
        one.d2h(img, d_img)
        d_o8 = one.dev_alloc(w * h)
        one.energy_image_dev(d_out, w, w, h, d_o8, w)
        want8 = np.empty((h, w), np.uint8)
        one.d2h(want8, d_o8)
        m = dc.Multi(None, blocksize=b)
        got, got8 = m.energy_bands(img, want_image=True)
        m.bands_create(w, h, ch)
        m.bands_synth(seed)
        m.bands_energy(sync=True)
        got2 = m.bands_download()
        same = np.array_equal(got, want) and np.array_equal(got2, want) and np.array_equal(got8, want8)
        ok &= same
        print("dctc_multi b=%d devices=%d bands==single-GPU (map, resident map, energy image): %s" % (b, m.n, same), flush=True)
        if b == 8:
            imgs = np.stack([img[:1080, :1920]] * 3 + [img[100:1180, 200:2120]] * 2)
            imgs = np.ascontiguousarray(imgs)
            same = np.array_equal(m.energy_batch(imgs), one.energy_batch(imgs))
            ok &= same
            print("dctc_multi_energy_batch (5 frames round-robin over %d devices) == one device: %s" % (m.n, same), flush=True)
        m.close()
        one.close()
    print("MULTI_CHECK single-process", "PASS" if ok else "FAIL", "(%d devices)" % n, flush=True)
    return ok


def one_process_per_gpu():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    name = "check_%s_%s" % (os.environ.get("MASTER_PORT", "0"), os.environ.get("TORCHELASTIC_RUN_ID", "x"))
    ok = True
    w, h, ch, seed = 4096, 2051, 3, 77
    for b in (8, 16, 4, 2):
        ctx = dc.Context(local, blocksize=b)
        r = dc.BandRunner(ctx, "%s_b%d" % (name, b), rank, world, w, h, ch)
        r.synth(seed)
        r.connect()
        r.step(sync=True)
        band = r.fetch()
        im8 = r.energy_image()
        # gather through the C rendezvous as well (no framework on the data path)
        blobs = dc.rendezvous_allgather("%s_g%d" % (name, b), rank, world, np.int32([r.y0, r.band_rows]).tobytes())
        np.save("/dev/shm/%s_b%d_r%d.npy" % (name, b, rank), band)
        np.save("/dev/shm/%s_b%d_i%d.npy" % (name, b, rank), im8)
        r.barrier()
        if rank == 0:
            full = np.concatenate([np.load("/dev/shm/%s_b%d_r%d.npy" % (name, b, q)) for q in range(world)], 0)
            full8 = np.concatenate([np.load("/dev/shm/%s_b%d_i%d.npy" % (name, b, q)) for q in range(world)], 0)
            d_img = ctx.dev_alloc(w * h * ch)
            d_out = ctx.dev_alloc(w * h * 4)
            d_o8 = ctx.dev_alloc(w * h)
            ctx.synth_fill_dev(d_img, 1, 0, w, h, ch, w * ch, seed, 0)
            ctx.energy_batch_dev(d_img, 1, 0, w, h, ch, w * ch, d_out, 0, w, sync=True)
            ctx.energy_image_dev(d_out, w, w, h, d_o8, w)
            want = np.empty((h, w), np.float32)
            want8 = np.empty((h, w), np.uint8)
            ctx.d2h(want, d_out)
            ctx.d2h(want8, d_o8)
            same = np.array_equal(full, want) and np.array_equal(full8, want8)
            ok &= same
            rows = [tuple(np.frombuffer(x, np.int32)) for x in blobs]
            print("band runners b=%d world=%d bands %s == single GPU (map, energy image): %s" % (b, world, rows, same), flush=True)
            for p in (d_img, d_out, d_o8):
                ctx.dev_free(p)
        r.barrier()
        os.remove("/dev/shm/%s_b%d_r%d.npy" % (name, b, rank))
        os.remove("/dev/shm/%s_b%d_i%d.npy" % (name, b, rank))
        r.close()
        ctx.close()
    if rank == 0:
        print("MULTI_CHECK per-process", "PASS" if ok else "FAIL", "(world %d)" % world, flush=True)
    return ok


if __name__ == "__main__":
    good = one_process_per_gpu() if "RANK" in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1 else single_process()
    sys.exit(0 if good else 1)
