"""Row-band sharding of one tall image over the GPUs of a box, one process per GPU (BASELINE config 5) -- the round-1
Python plumbing.  The product path is the C layer now (csrc/dctc_multi.cu: dctc_band_runner_* / dctc_multi_*, mirrored
by dct_carver_b200.BandRunner / Multi); this module stays for the torch.distributed "exchange" mode and its gloo tests.

Every output pixel needs blocksize/2-1 rows above and blocksize/2 rows below (window offsets -b/2+1..b/2,
/root/reference/src/render.c:146-147), so a band needs that many rows from each neighbour; the image's own
top/bottom edges replicate (render.c:122-132).  Two ways to get the halo:
  * "peer": the neighbour's band buffer is mapped through CUDA IPC and the K1 kernel loads the halo rows over
    NVLink itself (dctc_energy_band_dev with d_top/d_bot in peer memory) — no exchange step, no staging;
  * "exchange": classic neighbour send/recv of the halo rows into local buffers (torch.distributed isend/irecv,
    NCCL on GPUs, gloo on CPU for the tests), then the kernel reads local halos.
This module is plumbing only (partitioning + handles + send/recv); all compute is in libdctc.so.
"""
import numpy as np


def band_bounds(h, world):
    """Rows [y0, y1) owned by each rank: contiguous, near-equal, every row owned exactly once."""
    return [(h * r // world, h * (r + 1) // world) for r in range(world)]


def halo_rows(blocksize):
    """(rows needed above, rows needed below)."""
    return blocksize // 2 - 1, blocksize // 2


def exchange_halos(dist, rank, world, band, top_need, bot_need):
    """band: torch tensor [rows, ...] on any device.  Returns (top_halo or None, bot_halo or None): the last
    `top_need` rows of rank-1's band and the first `bot_need` rows of rank+1's band."""
    import torch
    ops, top, bot = [], None, None
    rows = band.shape[0]
    if top_need > rows or bot_need > rows:
        raise ValueError("band thinner than the halo")
    if rank > 0:
        if top_need > 0:
            top = torch.empty((top_need,) + tuple(band.shape[1:]), dtype=band.dtype, device=band.device)
            ops.append(dist.P2POp(dist.irecv, top, rank - 1))
        if bot_need > 0:
            ops.append(dist.P2POp(dist.isend, band[:bot_need].contiguous(), rank - 1))
    if rank < world - 1:
        if bot_need > 0:
            bot = torch.empty((bot_need,) + tuple(band.shape[1:]), dtype=band.dtype, device=band.device)
            ops.append(dist.P2POp(dist.irecv, bot, rank + 1))
        if top_need > 0:
            ops.append(dist.P2POp(dist.isend, band[rows - top_need:].contiguous(), rank + 1))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return top, bot


class BandRunner:
    """One rank's band of a w x h image with synthetic content generated on the device."""

    def __init__(self, ctx, dist, rank, world, w, h, ch, seed, mode="peer"):
        self.ctx, self.dist, self.rank, self.world = ctx, dist, rank, world
        self.w, self.h, self.ch = w, h, ch
        self.y0, self.y1 = band_bounds(h, world)[rank]
        self.band_rows = self.y1 - self.y0
        self.pitch = w * ch
        self.top_need, self.bot_need = halo_rows(ctx.blocksize)
        self.mode = mode if world > 1 else "single"
        # every band must be tall enough to serve its neighbours' halos (a kernel reads at most one band away)
        rows_all = [y1 - y0 for y0, y1 in band_bounds(h, world)]
        if world > 1 and min(rows_all) < max(self.top_need, self.bot_need):
            raise ValueError("band thinner than the halo: %d rows, %d needed" % (min(rows_all), max(self.top_need, self.bot_need)))
        self.d_band = ctx.dev_alloc(self.band_rows * self.pitch)
        self.d_out = ctx.dev_alloc(self.band_rows * w * 4)
        # content = rows y0..y1 of the virtual image (y_offset keeps bands consistent across ranks)
        for r0 in range(0, self.band_rows, 16384):
            n = min(16384, self.band_rows - r0)
            ctx.synth_fill_dev(self.d_band + r0 * self.pitch, 1, 0, w, n, ch, self.pitch, seed, 0, 0, self.y0 + r0)
        ctx.sync()
        self.d_top = self.d_bot = None
        self.top_rows = self.bot_rows = 0
        self._peers = []
        if self.mode == "peer":
            self._map_peers()
        elif self.mode == "exchange":
            self._alloc_exchange()

    def _map_peers(self):
        import torch
        dist = self.dist
        mine = torch.tensor(list(self.ctx.ipc_export(self.d_band)), dtype=torch.uint8, device="cuda")
        rows = torch.tensor([self.band_rows], dtype=torch.int64, device="cuda")
        hs = [torch.empty_like(mine) for _ in range(self.world)]
        rs = [torch.empty_like(rows) for _ in range(self.world)]
        dist.all_gather(hs, mine)
        dist.all_gather(rs, rows)
        if self.rank > 0 and self.top_need > 0:
            p = self.ctx.ipc_open(bytes(hs[self.rank - 1].cpu().numpy().tolist()))
            self._peers.append(p)
            self.top_rows = self.top_need
            self.d_top = p + (int(rs[self.rank - 1].item()) - self.top_rows) * self.pitch
        if self.rank < self.world - 1 and self.bot_need > 0:
            p = self.ctx.ipc_open(bytes(hs[self.rank + 1].cpu().numpy().tolist()))
            self._peers.append(p)
            self.bot_rows = self.bot_need
            self.d_bot = p
        dist.barrier()

    def _alloc_exchange(self):
        import torch
        self.t_band = None  # exchange mode stages through torch tensors created lazily in step()
        self.t_top = (torch.empty((self.top_need, self.pitch), dtype=torch.uint8, device="cuda")
                      if self.rank > 0 and self.top_need > 0 else None)
        self.t_bot = (torch.empty((self.bot_need, self.pitch), dtype=torch.uint8, device="cuda")
                      if self.rank < self.world - 1 and self.bot_need > 0 else None)

    def step(self):
        c = self.ctx
        if self.mode == "exchange":
            self._exchange()
        c.energy_band_dev(self.d_band, self.w, self.band_rows, self.ch, self.pitch, self.d_top, self.top_rows,
                          self.pitch, self.d_bot, self.bot_rows, self.pitch, self.d_out, self.w, band_y0=self.y0)

    def _exchange(self):
        import torch
        dist = self.dist
        c = self.ctx
        ops = []
        # wrap raw device rows as torch tensors without copying
        def view(ptr, rows):
            arr = _DevArray(ptr, (rows, self.pitch))
            return torch.as_tensor(arr, device="cuda")
        if self.rank > 0:
            if self.top_need > 0:
                ops.append(dist.P2POp(dist.irecv, self.t_top, self.rank - 1))
            if self.bot_need > 0:
                ops.append(dist.P2POp(dist.isend, view(self.d_band, self.bot_need), self.rank - 1))
        if self.rank < self.world - 1:
            if self.bot_need > 0:
                ops.append(dist.P2POp(dist.irecv, self.t_bot, self.rank + 1))
            if self.top_need > 0:
                ops.append(dist.P2POp(dist.isend, view(self.d_band + (self.band_rows - self.top_need) * self.pitch, self.top_need), self.rank + 1))
        for req in (dist.batch_isend_irecv(ops) if ops else []):
            req.wait()
        torch.cuda.current_stream().synchronize()
        if self.t_top is not None:
            self.d_top, self.top_rows = self.t_top.data_ptr(), self.top_need
        if self.t_bot is not None:
            self.d_bot, self.bot_rows = self.t_bot.data_ptr(), self.bot_need

    def fetch(self):
        out = np.empty((self.band_rows, self.w), np.float32)
        self.ctx.d2h(out, self.d_out)
        return out

    def close(self):
        if self.dist is not None and self.world > 1:
            self.ctx.sync()
            self.dist.barrier()        # no rank frees its band while a neighbour's kernel may still be reading it
        for p in self._peers:
            self.ctx.ipc_close(p)
        self._peers = []
        self.ctx.dev_free(self.d_band)
        self.ctx.dev_free(self.d_out)


class _DevArray:
    """__cuda_array_interface__ shim so torch can view raw device memory owned by libdctc."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "|u1", "data": (int(ptr), False), "version": 2}
