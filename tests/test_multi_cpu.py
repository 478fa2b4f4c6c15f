"""CPU tests of the C multi-GPU host layer (csrc/dctc_multi.cu): band geometry and the shared-memory rendezvous that
the one-process-per-GPU band runners use to exchange their CUDA IPC handles.  No compute calls (no GPU here)."""
import multiprocessing as mp
import os

import pytest

import dct_carver_b200 as dc


def test_band_plan_covers_every_row_once_and_matches_window_offsets():
    for h in (8, 100, 2160, 32768):
        for world in (1, 2, 3, 4, 8):
            for b in (2, 4, 8, 16):
                if world > 1 and h // world < b // 2:
                    continue
                plans = [dc.band_plan(h, world, r, b) for r in range(world)]
                assert plans[0][0] == 0 and plans[-1][0] + plans[-1][1] == h
                assert all(plans[i][0] + plans[i][1] == plans[i + 1][0] for i in range(world - 1))
                sizes = [p[1] for p in plans]
                assert max(sizes) - min(sizes) <= 1
                # window offsets -b/2+1 .. b/2 (src/render.c:146-147): b/2-1 rows above, b/2 below; none at the image edges
                assert plans[0][2] == 0 and plans[-1][3] == 0
                for r in range(world):
                    assert plans[r][2] == (b // 2 - 1 if r > 0 else 0)
                    assert plans[r][3] == (b // 2 if r < world - 1 else 0)


def test_band_plan_rejects_bands_thinner_than_the_halo():
    L = dc.lib()
    assert L.dctc_band_plan(12, 8, 0, 16, None, None, None, None) == dc.ERR_INVALID    # 1-2 rows per band, halo 8
    assert L.dctc_band_plan(12, 8, 0, 2, None, None, None, None) == dc.OK
    assert L.dctc_band_plan(100, 4, 4, 8, None, None, None, None) == dc.ERR_INVALID     # rank out of range
    assert L.dctc_band_plan(100, 4, 0, 5, None, None, None, None) == dc.ERR_BLOCKSIZE


def _rv_worker(name, rank, world, q):
    try:
        blob = bytes([rank + 1]) * 64 + rank.to_bytes(4, "little")
        got = dc.rendezvous_allgather(name, rank, world, blob, timeout_ms=30000)
        second = dc.rendezvous_allgather(name + "_b", rank, world, bytes([9 - rank]), timeout_ms=30000)   # a second round
        q.put((rank, got, second))
    except Exception as e:   # pragma: no cover
        q.put((rank, repr(e), None))


@pytest.mark.parametrize("world", [2, 3])
def test_rendezvous_allgather_between_processes(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    name = "pytest_%d_%d" % (os.getpid(), world)
    procs = [ctx.Process(target=_rv_worker, args=(name, r, world, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [bytes([r + 1]) * 64 + r.to_bytes(4, "little") for r in range(world)]
    for rank, got, second in res:
        assert got == want, (rank, got)
        assert second == [bytes([9 - r]) for r in range(world)]
    assert not os.path.exists("/dev/shm/dctc_" + name)          # the last rank to leave unlinks the segment


def test_rendezvous_times_out_instead_of_hanging():
    with pytest.raises(dc.DctcError) as e:
        dc.rendezvous_allgather("pytest_lonely_%d" % os.getpid(), 0, 2, b"x", timeout_ms=200)
    assert e.value.status == dc.ERR_STATE
    assert not os.path.exists("/dev/shm/dctc_pytest_lonely_%d" % os.getpid())
