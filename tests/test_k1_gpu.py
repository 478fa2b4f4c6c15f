"""GPU parity tests of K1 (full energy map), through the C ABI, against the oracle / compiled reference / golden
vectors.  Tolerance (FP32 CUDA path vs the double-precision reference): |err| <= 1e-6 + 4e-6*|ref|."""
import ctypes as C
import os

import numpy as np
import pytest

import dct_carver_b200 as dc
import oracle_lib as ol

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "energy_golden.npz"))
N_GOLD = sum(1 for k in GOLD.files if k.startswith("img_"))
KERNELS = [dc.KERNEL_FP32_TILE, dc.KERNEL_FP32_MARCH, dc.KERNEL_TC_SPLIT]


def check_with_flips(got, img, b, e, t, max_flip_frac=2e-3):
    """Non-uniform weights: a near-tie between an edge atom and a texture atom may resolve differently in FP32
    (SURVEY section 7 'class flips').  Such pixels must equal the other class's product and be rare."""
    want, cls = ol.oracle_energy(img, b, e, t, want_class=True)
    got64, want64 = got.astype(np.float64), want.astype(np.float64)
    ok = np.abs(got64 - want64) <= ol.ABS_TOL + ol.REL_TOL * np.abs(want64)
    if e == t or ok.all():
        assert ok.all(), ol.parity(got, want)
        return 0
    w_this = np.where(cls == 1, e, t).astype(np.float64)
    w_other = np.where(cls == 1, t, e).astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        alt = np.where(w_this > 0, want64 / w_this * w_other, np.nan)
    flip_ok = np.abs(got64 - alt) <= ol.ABS_TOL + 1e-5 * np.abs(alt)
    assert (ok | flip_ok).all(), "mismatch that is not a class flip: %d" % (~(ok | flip_ok)).sum()
    flips = int((~ok).sum())
    assert flips <= max(2, max_flip_frac * got.size), flips
    return flips


@pytest.fixture(scope="module")
def ctx():
    c = dc.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("i", range(N_GOLD))
def test_golden_vectors(ctx, i, kernel):
    img = GOLD["img_%02d" % i]
    want = GOLD["en_%02d" % i]
    b = int(GOLD["meta_%02d" % i][0])
    e, t = (float(v) for v in GOLD["wts_%02d" % i])
    ctx.set_params(b, e, t)
    ctx.set_kernel(kernel if b == 8 else dc.KERNEL_AUTO)
    got = ctx.energy_full(img)
    if e == t:
        ol.assert_parity(got, want)
    else:
        check_with_flips(got, img, b, e, t)


@pytest.mark.parametrize("b", [2, 4, 8, 16])
@pytest.mark.parametrize("wts", [(0.5, 0.5), (0.8, 0.2), (1.0, 0.0)])
@pytest.mark.parametrize("case", [(0, 3, 200, 150), (0, 1, 131, 67), (3, 3, 96, 130), (1, 4, 70, 41), (2, 2, 65, 33),
                                  (0, 3, 1, 1), (0, 1, 3, 2), (0, 3, 7, 40), (0, 1, 64, 32), (0, 3, 65, 33)])
def test_parity_vs_oracle(ctx, b, wts, case):
    pattern, ch, w, h = case
    img = ol.synth_image(w, h, ch, 1000 + b, pattern)
    ctx.set_params(b, *wts)
    ctx.set_kernel(dc.KERNEL_AUTO)
    got = ctx.energy_full(img)
    assert got.shape == (h, w)
    check_with_flips(got, img, b, *wts)


@pytest.mark.parametrize("b", [2, 4, 8, 16])
def test_512_grey_config1_against_best_checker(ctx, b):
    """BASELINE config 1: 512x512 8-bit grey, full map; checker = the compiled reference when it travelled."""
    img = ol.synth_image(512, 512, 1, 0xD0C7CA12, 0)
    ctx.set_params(b, 0.5, 0.5)
    ctx.set_kernel(dc.KERNEL_AUTO)
    got = ctx.energy_full(img)
    ol.assert_parity(got, ol.best_energy(img, b, 0.5, 0.5))


def test_pitch_and_subimage(ctx):
    """Row pitch larger than w*channels (the carver hands us sub-rectangles of bigger buffers)."""
    big = ol.synth_image(160, 90, 3, 5, 0)
    sub = np.ascontiguousarray(big[10:70, 20:140])
    ctx.set_params(8, 0.5, 0.5)
    want = ctx.energy_full(sub)
    out = np.empty((60, 120), np.float32)
    view = big[10:70, 20:140]
    rc = dc.lib().dctc_energy_full(ctx.handle, C.c_void_p(view.ctypes.data), 120, 60, 3, big.strides[0],
                                   C.c_void_p(out.ctypes.data))
    assert rc == 0
    assert np.array_equal(out, want)


def test_error_behaviour(ctx):
    L = dc.lib()
    p = dc.EnergyParameters(edges=0.5, textures=0.5, blocksize=3)
    assert L.dctc_set_params(ctx.handle, C.byref(p)) == dc.ERR_BLOCKSIZE     # dct.c:89-92 -> error()
    img = np.zeros((4, 4, 3), np.uint8)
    out = np.zeros((4, 4), np.float32)
    assert L.dctc_energy_full(ctx.handle, img.ctypes.data, 4, 4, 5, 20, out.ctypes.data) == dc.ERR_INVALID
    assert L.dctc_energy_full(ctx.handle, img.ctypes.data, 4, 4, 3, 8, out.ctypes.data) == dc.ERR_INVALID
    assert L.dctc_energy_full(ctx.handle, img.ctypes.data, 0, 4, 3, 12, out.ctypes.data) == dc.ERR_INVALID
    assert L.dctc_energy_full(ctx.handle, None, 4, 4, 3, 12, out.ctypes.data) == dc.ERR_INVALID
    ctx.set_params(8, 0.5, 0.5)


def test_batch_equals_single_frames(ctx):
    n, h, w, ch = 5, 45, 77, 3
    imgs = np.stack([ol.synth_image(w, h, ch, 42, 0, frame=f) for f in range(n)])
    ctx.set_params(8, 0.5, 0.5)
    got = ctx.energy_batch(imgs)
    for f in range(n):
        assert np.array_equal(got[f], ctx.energy_full(imgs[f]))


def test_device_synth_matches_host_generator(ctx):
    w, h, ch, n = 130, 37, 3, 3
    d = ctx.dev_alloc(n * w * h * ch)
    for pattern in range(4):
        ctx.synth_fill_dev(d, n, w * h * ch, w, h, ch, w * ch, 99, pattern, first_frame=2, y_offset=11)
        out = np.empty((n, h, w, ch), np.uint8)
        ctx.d2h(out, d)
        for f in range(n):
            assert np.array_equal(out[f], ol.synth_image(w, h, ch, 99, pattern, frame=2 + f, y_offset=11)), pattern
    ctx.dev_free(d)


def _band_run(ctx, img, bounds, b):
    """Energy of a tall image computed band by band with device-resident halos (config 5 on one GPU)."""
    h, w, ch = img.shape
    pitch = w * ch
    d_img = ctx.dev_alloc(img.nbytes)
    d_out = ctx.dev_alloc(h * w * 4)
    ctx.h2d(d_img, img)
    r_top, r_bot = b // 2 - 1, b // 2
    for (y0, y1) in bounds:
        t = min(r_top, y0)
        bt = min(r_bot, h - y1)
        ctx.energy_band_dev(d_img + y0 * pitch, w, y1 - y0, ch, pitch,
                            d_img + (y0 - t) * pitch if t else None, t, pitch,
                            d_img + y1 * pitch if bt else None, bt, pitch,
                            d_out + y0 * w * 4, w)
    out = np.empty((h, w), np.float32)
    ctx.d2h(out, d_out)
    ctx.dev_free(d_img)
    ctx.dev_free(d_out)
    return out


@pytest.mark.parametrize("b", [2, 4, 8, 16])
def test_row_bands_with_halo_equal_full_image(ctx, b):
    """Row bands + halos reproduce the full map BIT-exactly when one kernel serves both (the FP32 kernels here; rows
    of a 150-px RGB image are not 16-byte aligned, which is outside the tensor-core kernel's fast path)."""
    img = ol.synth_image(150, 101, 3, 7, 0)
    ctx.set_params(b, 0.5, 0.5)
    ctx.set_kernel(dc.KERNEL_FP32_MARCH)
    full = ctx.energy_full(img)
    got = _band_run(ctx, img, [(0, 13), (13, 50), (50, 51), (51, 101)], b)
    ctx.set_kernel(dc.KERNEL_AUTO)
    assert np.array_equal(got, full)


@pytest.mark.parametrize("ch,w", [(3, 160), (1, 208)])
def test_row_bands_tensor_core_kernel(ctx, ch, w):
    """Same property through the tensor-core kernel (16-byte aligned rows, so every band stays on its fast path):
    the band decomposition changes which ring slot / MMA step a row falls into, and the result must not depend on it."""
    img = ol.synth_image(w, 131, ch, 17, 0)
    ctx.set_params(8, 0.5, 0.5)
    ctx.set_kernel(dc.KERNEL_TC_SPLIT)
    full = ctx.energy_full(img)
    got = _band_run(ctx, img, [(0, 13), (13, 50), (50, 51), (51, 131)], 8)
    ctx.set_kernel(dc.KERNEL_AUTO)
    ol.assert_parity(got, ol.best_energy(img, 8, 0.5, 0.5))
    assert np.array_equal(got, full)


@pytest.mark.parametrize("b", [2, 4])
@pytest.mark.parametrize("wts", [(0.5, 0.5), (0.8, 0.2)])
@pytest.mark.parametrize("case", [(0, 3, 160, 131), (0, 1, 208, 77), (3, 3, 640, 300), (1, 3, 128, 8), (2, 3, 144, 50),
                                  (0, 3, 16, 5), (0, 1, 16, 3), (0, 3, 272, 9), (0, 1, 1008, 40)])
def test_stream_kernel_small_blocks(ctx, b, wts, case):
    """Block sizes 2 and 4 on 16-byte aligned rows take the streaming register-march kernel (dctc_k1_small.cu): parity
    with the oracle; grey maps are bit-identical to the tile kernel's (same FP32 operation order), RGB maps agree with
    it within the tolerance (the streaming kernel computes luma as an exact integer, the tile kernel as an FMA chain)."""
    pattern, ch, w, h = case
    assert (w * ch) % 16 == 0
    img = ol.synth_image(w, h, ch, 2000 + b, pattern)
    ctx.set_params(b, *wts)
    ctx.set_kernel(dc.KERNEL_AUTO)
    got = ctx.energy_full(img)
    ctx.set_kernel(dc.KERNEL_FP32_TILE)
    tile = ctx.energy_full(img)
    ctx.set_kernel(dc.KERNEL_AUTO)
    if ch == 1:
        assert np.array_equal(got.view(np.uint32), tile.view(np.uint32))
    elif wts[0] == wts[1]:
        ol.assert_parity(got, tile, rel=2e-6, abs_=1e-6)
    check_with_flips(got, img, b, *wts)


@pytest.mark.parametrize("b", [8, 16])
def test_back_to_back_launches_are_ordered(ctx, b):
    """The tensor-core launches are chained by programmatic dependent launches (the ramp of a launch overlaps the tail of
    its predecessor).  A launch whose INPUT is the previous launch's OUTPUT (the float map read as a grey image four
    times as wide) must still see the complete map: same result as with a host synchronisation in between."""
    w, h = 512, 270
    img = ol.synth_image(w, h, 1, 4242, 0)
    ctx.set_params(b, 0.5, 0.5)
    ctx.set_kernel(dc.KERNEL_AUTO)
    d_img = ctx.dev_alloc(w * h)
    ctx.h2d(d_img, img)
    d_mid = ctx.dev_alloc(w * h * 4)
    d_out = ctx.dev_alloc(4 * w * h * 4)
    res = []
    for sync_between in (False, True):
        ctx.h2d(d_mid, np.zeros(w * h, np.float32))
        for _ in range(3):                                    # keep the stream busy so that launches really follow each other
            ctx.energy_batch_dev(d_img, 1, 0, w, h, 1, w, d_mid, 0, w, sync=sync_between)
            ctx.energy_batch_dev(d_mid, 1, 0, 4 * w, h, 1, 4 * w, d_out, 0, 4 * w, sync=sync_between)
        ctx.sync()
        got = np.empty((h, 4 * w), np.float32)
        ctx.d2h(got, d_out)
        res.append(got)
    for d in (d_img, d_mid, d_out):
        ctx.dev_free(d)
    assert np.array_equal(res[0].view(np.uint32), res[1].view(np.uint32))
    assert float(np.abs(res[0]).max()) > 0.0


@pytest.mark.parametrize("b", [8, 16, 4])
def test_batch_with_row_and_frame_padding(ctx, b):
    """Frames with a row pitch larger than the row and a frame stride larger than the frame (both 16-byte multiples, so
    every kernel stays on its fast path; block size 8 stages its raw tiles through a 3-D tensor map built from exactly
    these strides): each frame of the batch equals the same frame computed alone from a tight buffer."""
    n, w, h, ch = 3, 200, 77, 3
    pitch = (w * ch + 15) // 16 * 16 + 32
    fs = h * pitch + 4096
    buf = np.full(n * fs, 0xEE, np.uint8)                      # padding bytes must not matter
    frames = [ol.synth_image(w, h, ch, 700 + b, 0, frame=f) for f in range(n)]
    for f in range(n):
        v = buf[f * fs:f * fs + h * pitch].reshape(h, pitch)
        v[:, :w * ch] = frames[f].reshape(h, w * ch)
    ctx.set_params(b, 0.5, 0.5)
    ctx.set_kernel(dc.KERNEL_AUTO)
    d_img = ctx.dev_alloc(buf.nbytes)
    ctx.h2d(d_img, buf)
    ofs = w * h + 64
    d_out = ctx.dev_alloc(n * ofs * 4)
    ctx.energy_batch_dev(d_img, n, fs, w, h, ch, pitch, d_out, ofs, w, sync=True)
    got = np.empty(n * ofs, np.float32)
    ctx.d2h(got, d_out)
    ctx.dev_free(d_img)
    ctx.dev_free(d_out)
    for f in range(n):
        want = ctx.energy_full(frames[f])
        assert np.array_equal(got[f * ofs:f * ofs + w * h].reshape(h, w).view(np.uint32), want.view(np.uint32)), f
        ol.assert_parity(want, ol.best_energy(frames[f], b, 0.5, 0.5))


@pytest.mark.parametrize("b", [2, 4])
@pytest.mark.parametrize("w,out_pitch,odd_base", [(272, 275, 0), (272, 272, 1), (300, 300, 0), (259, 260, 0)])
def test_stream_kernel_output_layouts(ctx, b, w, out_pitch, odd_base):
    """The streaming kernel writes one 64-bit value per thread and row when the output rows are 8-byte aligned, and two
    predicated 32-bit values otherwise (odd output pitch, output pointer on an odd float, last column pair of an odd
    width, strips that stick out of the image): every layout gives the same map, and nothing outside it is written."""
    ch, h, n = 1, 21, 2
    pitch = (w * ch + 15) // 16 * 16
    imgs = np.zeros((n, h, pitch), np.uint8)
    for f in range(n):
        imgs[f, :, :w] = ol.synth_image(w, h, ch, 500 + b, 0, frame=f).reshape(h, w)
    ctx.set_params(b, 0.5, 0.5)
    d_img = ctx.dev_alloc(imgs.nbytes)
    ctx.h2d(d_img, imgs)
    ofs = h * out_pitch + 3                      # floats between output frames (odd or even as it comes)
    total = n * ofs + 8
    d_out = ctx.dev_alloc(total * 4)
    sentinel = np.full(total, -7.0, np.float32)
    ctx.h2d(d_out, sentinel)
    ctx.set_kernel(dc.KERNEL_FP32_STREAM)
    ctx.energy_batch_dev(d_img, n, h * pitch, w, h, ch, pitch, d_out + 4 * odd_base, ofs, out_pitch, sync=True)
    ctx.set_kernel(dc.KERNEL_AUTO)
    got = np.empty(total, np.float32)
    ctx.d2h(got, d_out)
    ctx.dev_free(d_img)
    ctx.dev_free(d_out)
    written = np.zeros(total, bool)
    for f in range(n):
        ctx.set_kernel(dc.KERNEL_FP32_TILE)
        want = ctx.energy_full(np.ascontiguousarray(imgs[f, :, :w]).reshape(h, w, 1))
        ctx.set_kernel(dc.KERNEL_AUTO)
        for y in range(h):
            o = odd_base + f * ofs + y * out_pitch
            assert np.array_equal(got[o:o + w].view(np.uint32), want[y].view(np.uint32)), (f, y)
            written[o:o + w] = True
    assert np.all(got[~written] == -7.0)


@pytest.mark.parametrize("b", [2, 4, 8])
def test_stream_kernel_selector(ctx, b):
    """dctc_set_kernel(DCTC_KERNEL_FP32_STREAM) names the streaming kernel of block sizes 2 and 4 explicitly (it is also what
    DCTC_KERNEL_AUTO picks for them); block size 8 has no streaming kernel and takes its FP32 register-march kernel."""
    img = ol.synth_image(272, 37, 3, 99 + b, 0)
    ctx.set_params(b, 0.5, 0.5)
    ctx.set_kernel(dc.KERNEL_FP32_STREAM)
    got = ctx.energy_full(img)
    ctx.set_kernel(dc.KERNEL_AUTO if b != 8 else dc.KERNEL_FP32_MARCH)
    want = ctx.energy_full(img)
    ctx.set_kernel(dc.KERNEL_AUTO)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    ol.assert_parity(got, ol.best_energy(img, b, 0.5, 0.5))


@pytest.mark.parametrize("b", [2, 4])
@pytest.mark.parametrize("ch,w", [(3, 160), (1, 208)])
def test_row_bands_stream_kernel(ctx, b, ch, w):
    """Row bands + halos through the streaming kernel (aligned rows keep every band on its fast path) are bit-equal
    to the full map."""
    img = ol.synth_image(w, 131, ch, 23, 0)
    ctx.set_params(b, 0.5, 0.5)
    ctx.set_kernel(dc.KERNEL_AUTO)
    full = ctx.energy_full(img)
    got = _band_run(ctx, img, [(0, 13), (13, 50), (50, 51), (51, 131)], b)
    ol.assert_parity(got, ol.best_energy(img, b, 0.5, 0.5))
    assert np.array_equal(got, full)


def test_full_size_4k_properties(ctx):
    """BASELINE config 2 (3840x2160 RGB) through size-independent properties: (1) random crops re-evaluated by the
    oracle on crop+halo agree in the crop interior (the operator is local); (2) row-band decomposition is bit-equal;
    (3) the checksum of per-row checksums is reproducible across two runs."""
    w, h, ch = 3840, 2160, 3
    ctx.set_params(8, 0.5, 0.5)
    d_img = ctx.dev_alloc(w * h * ch)
    ctx.synth_fill_dev(d_img, 1, 0, w, h, ch, w * ch, 0xD0C7CA13, 0)
    img = np.empty((h, w, ch), np.uint8)
    ctx.d2h(img, d_img)
    assert np.array_equal(img[1000:1003, 2000:2005], ol.synth_image(w, h, ch, 0xD0C7CA13, 0)[1000:1003, 2000:2005])
    d_out = ctx.dev_alloc(w * h * 4)
    ctx.energy_batch_dev(d_img, 1, 0, w, h, ch, w * ch, d_out, 0, w, sync=True)
    en = np.empty((h, w), np.float32)
    ctx.d2h(en, d_out)
    rng = np.random.default_rng(3)
    for _ in range(12):
        y0 = int(rng.integers(8, h - 56))
        x0 = int(rng.integers(8, w - 56))
        crop = img[y0 - 8:y0 + 56, x0 - 8:x0 + 56]
        want = ol.best_energy(crop, 8, 0.5, 0.5)[8:-8, 8:-8]
        ol.assert_parity(en[y0:y0 + 48, x0:x0 + 48], want)
    # corners: the edge replication at full size (window offsets -3..+4, so a 48-px crop covers 40 px exactly)
    tl = ol.best_energy(np.ascontiguousarray(img[:48, :48]), 8, 0.5, 0.5)[:40, :40]
    ol.assert_parity(en[:40, :40], tl)
    br = ol.best_energy(np.ascontiguousarray(img[h - 48:, w - 48:]), 8, 0.5, 0.5)[8:, 8:]
    ol.assert_parity(en[h - 40:, w - 40:], br)
    tr = ol.best_energy(np.ascontiguousarray(img[:48, w - 48:]), 8, 0.5, 0.5)[:40, 8:]
    ol.assert_parity(en[:40, w - 40:], tr)
    got = _band_run(ctx, img, [(0, 270), (270, 1080), (1080, 1081), (1081, 2160)], 8)
    assert np.array_equal(got, en)
    ctx.energy_batch_dev(d_img, 1, 0, w, h, ch, w * ch, d_out, 0, w, sync=True)
    en2 = np.empty((h, w), np.float32)
    ctx.d2h(en2, d_out)
    assert np.float64(en2.sum(axis=1, dtype=np.float64)).sum() == np.float64(en.sum(axis=1, dtype=np.float64)).sum()
    ctx.dev_free(d_img)
    ctx.dev_free(d_out)


# ---- block size 16 on the tensor cores (dctc_k1_tc16.cu) -------------------------------------------------------------

@pytest.mark.parametrize("wts", [(0.5, 0.5), (0.8, 0.2), (1.0, 0.0)])
@pytest.mark.parametrize("case", [(0, 3, 64, 16), (0, 3, 208, 150), (0, 1, 144, 67), (3, 3, 96, 130), (3, 3, 640, 300),
                                  (1, 3, 128, 48), (2, 3, 528, 77), (0, 1, 16, 5), (0, 3, 1, 1), (0, 3, 7, 40)])
def test_tc16_parity(ctx, wts, case):
    """Block size 16 through the tcgen05 kernel (16-byte aligned device rows: the host API stages them that way)
    against the compiled reference / oracle, and within tolerance of the FP32 tile kernel."""
    pattern, ch, w, h = case
    img = ol.synth_image(w, h, ch, 1600 + w, pattern)
    ctx.set_params(16, *wts)
    ctx.set_kernel(dc.KERNEL_AUTO)
    got = ctx.energy_full(img)
    ctx.set_kernel(dc.KERNEL_FP32_TILE)
    tile = ctx.energy_full(img)
    ctx.set_kernel(dc.KERNEL_AUTO)
    check_with_flips(got, img, 16, *wts)
    if wts[0] == wts[1]:
        assert np.abs(got.astype(np.float64) - tile).max() <= 2 * (ol.ABS_TOL + ol.REL_TOL * float(np.abs(tile).max()))


@pytest.mark.parametrize("ch,w", [(3, 160), (1, 208)])
def test_tc16_row_bands_bit_equal_with_band_origin(ctx, ch, w):
    """dctc_energy_band_dev_at: with the band's image row given, the 16-row accumulation steps of the tensor-core kernel
    are anchored to the image's row grid, so any cut into bands reproduces the full map BIT for bit (and without the
    origin it still agrees within the tolerance)."""
    h = 131
    img = ol.synth_image(w, h, ch, 1617, 0)
    ctx.set_params(16, 0.5, 0.5)
    ctx.set_kernel(dc.KERNEL_AUTO)
    full = ctx.energy_full(img)
    bounds = [(0, 13), (13, 50), (50, 51), (51, 99), (99, 131)]
    pitch = (w * ch + 15) & ~15
    buf = np.zeros((h, pitch), np.uint8)
    buf[:, :w * ch] = img.reshape(h, w * ch)
    d_img = ctx.dev_alloc(h * pitch)
    d_out = ctx.dev_alloc(h * w * 4)
    ctx.h2d(d_img, buf)
    res = []
    for with_origin in (True, False):
        for (y0, y1) in bounds:
            t, bt = min(7, y0), min(8, h - y1)
            ctx.energy_band_dev(d_img + y0 * pitch, w, y1 - y0, ch, pitch, d_img + (y0 - t) * pitch if t else None, t, pitch,
                                d_img + y1 * pitch if bt else None, bt, pitch, d_out + y0 * w * 4, w,
                                band_y0=y0 if with_origin else 0)
        out = np.empty((h, w), np.float32)
        ctx.d2h(out, d_out)
        res.append(out)
    ctx.dev_free(d_img)
    ctx.dev_free(d_out)
    assert np.array_equal(res[0], full)
    ol.assert_parity(res[1], ol.best_energy(img, 16, 0.5, 0.5))


def test_tc16_persistent_loop_many_items(ctx):
    """More work items than CTAs (400 small frames in one launch: every CTA re-enters the item loop and re-initialises
    its barriers), each frame equal to its single-frame launch."""
    n, w, h, ch = 400, 64, 48, 3
    imgs = np.stack([ol.synth_image(w, h, ch, 1616, 0, frame=f) for f in range(n)])
    ctx.set_params(16, 0.5, 0.5)
    ctx.set_kernel(dc.KERNEL_AUTO)
    d_in = ctx.dev_alloc(imgs.nbytes)
    d_out = ctx.dev_alloc(n * w * h * 4)
    ctx.h2d(d_in, imgs)
    ctx.energy_batch_dev(d_in, n, w * h * ch, w, h, ch, w * ch, d_out, w * h, w, sync=True)
    got = np.empty((n, h, w), np.float32)
    ctx.d2h(got, d_out)
    ctx.dev_free(d_in)
    ctx.dev_free(d_out)
    for f in (0, 1, 147, 148, 149, 295, 296, 399):
        assert np.array_equal(got[f], ctx.energy_full(imgs[f])), f
    ol.assert_parity(got[399], ol.best_energy(imgs[399], 16, 0.5, 0.5))
