"""CPU tests (gloo, world_size 2 and 3) of the N>1 row-band path: partitioning, halo exchange, edge handling.
Compute here is the oracle (the checker); on GPUs the same plumbing feeds dctc_energy_band_dev."""
import os
import socket

import numpy as np
import pytest

import oracle_lib as ol
from dct_carver_b200 import multigpu


def test_band_bounds_cover_every_row_once():
    for h in (1, 7, 100, 2160, 32768):
        for world in (1, 2, 3, 4, 8):
            b = multigpu.band_bounds(h, world)
            assert b[0][0] == 0 and b[-1][1] == h
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [y1 - y0 for y0, y1 in b]
            assert max(sizes) - min(sizes) <= 1


def test_halo_rows_match_window_offsets():
    # window offsets -b/2+1 .. b/2 (render.c:146-147)
    assert [multigpu.halo_rows(b) for b in (2, 4, 8, 16)] == [(0, 1), (1, 2), (3, 4), (7, 8)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, b, w, h, ch, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        y0, y1 = multigpu.band_bounds(h, world)[rank]
        img = ol.synth_image(w, h, ch, 2024, 0)           # every rank can regenerate the virtual image
        band = torch.from_numpy(np.ascontiguousarray(img[y0:y1]))
        tn, bn = multigpu.halo_rows(b)
        top, bot = multigpu.exchange_halos(dist, rank, world, band, tn, bn)
        parts = [t for t in (top, band, bot) if t is not None]
        ext = torch.cat(parts, 0).numpy()
        off = 0 if top is None else top.shape[0]
        # interior bands: halos make the clamp unreachable; edge bands: the image edge IS the buffer edge
        en = ol.oracle_energy(ext, b, 0.5, 0.5, nthreads=1)[off:off + (y1 - y0)]
        if top is not None:
            assert np.array_equal(top.numpy(), img[y0 - tn:y0])
        if bot is not None:
            assert np.array_equal(bot.numpy(), img[y1:y1 + bn])
        gathered = [None] * world
        dist.all_gather_object(gathered, (y0, en))
        if rank == 0:
            full = np.concatenate([e for _, e in sorted(gathered, key=lambda t: t[0])], 0)
            q.put(full)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,b", [(2, 8), (3, 8), (2, 16), (2, 2)])
def test_row_bands_with_halo_exchange_equal_full_image(world, b):
    import torch.multiprocessing as mp
    w, h, ch = 40, 45, 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, b, w, h, ch, q)) for r in range(world)]
    for p in procs:
        p.start()
    full = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = ol.oracle_energy(ol.synth_image(w, h, ch, 2024, 0), b, 0.5, 0.5)
    assert np.array_equal(full, want)
