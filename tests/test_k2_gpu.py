"""GPU tests of K2 (carver session: seam removal + incremental band energy) through the C ABI."""
import os
import numpy as np
import pytest

import dct_carver_b200 as dc
import oracle_lib as ol

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    # A carver session runs in one arithmetic (FP32 march + FP32 tile kernels, bit-identical to each other); the
    # "full recompute" references below therefore pin the FP32 march kernel instead of AUTO (= tensor cores).
    c = dc.Context(0, kernel=dc.KERNEL_FP32_MARCH)
    yield c
    c.close()


def random_seam(rng, w, h):
    x = int(rng.integers(0, w))
    s = np.empty(h, np.int32)
    for y in range(h):
        x = int(np.clip(x + rng.integers(-1, 2), 0, w - 1))
        s[y] = x
    return s


def carve_host(img, seam):
    h, w, ch = img.shape
    out = np.empty((h, w - 1, ch), img.dtype)
    for y in range(h):
        out[y] = np.delete(img[y], seam[y], axis=0)
    return out


@pytest.mark.parametrize("b", [2, 4, 8, 16])
@pytest.mark.parametrize("ch", [1, 3])
def test_incremental_equals_full_recompute(ctx, b, ch):
    """After every seam the device energy plane must be BIT-identical to K1 run from scratch on the carved image,
    and within tolerance of the oracle; the reported band must be liblqr's update_emap range."""
    rng = np.random.default_rng(b * 10 + ch)
    w, h = 97, 75
    img = ol.synth_image(w, h, ch, 11, 0)
    ctx.set_params(b, 0.5, 0.5)
    ctx.carver_load(img)
    assert np.array_equal(ctx.carver_energy(), ctx.energy_full(img))
    r = b // 2
    cur = img
    for k in range(12):
        seam = random_seam(rng, cur.shape[1], h) if k % 4 else np.full(h, (0 if k == 0 else cur.shape[1] - 1), np.int32)
        band, xmin, xmax = ctx.carve_and_update(seam)
        cur = carve_host(cur, seam)
        assert ctx.carver_size() == (cur.shape[1], h)
        assert np.array_equal(ctx.carver_image(), cur)
        en = ctx.carver_energy()
        full = ctx.energy_full(cur)
        assert np.array_equal(en, full), "seam %d" % k
        # band limits = SURVEY section 8b formula
        for y in range(h):
            lo = min(seam[max(0, y - r):y + r + 1]) - r
            hi = max(seam[max(0, y - r):y + r + 1]) + r - 1
            assert xmin[y] == max(lo, 0) and xmax[y] == min(hi, cur.shape[1] - 1)
        packed = np.concatenate([en[y, xmin[y]:xmax[y] + 1] for y in range(h)])
        assert np.array_equal(band, packed)
    ol.assert_parity(ctx.carver_energy(), ol.best_energy(cur, b, 0.5, 0.5))


def test_carver_state_errors(ctx):
    L = dc.lib()
    c2 = dc.Context(0)
    seam = np.zeros(4, np.int32)
    assert L.dctc_carve_and_update(c2.handle, seam.ctypes.data, None, None, None) == dc.ERR_STATE
    c2.carver_load(np.zeros((4, 6, 3), np.uint8))
    bad = np.array([0, 1, 6, 2], np.int32)
    assert L.dctc_carve_and_update(c2.handle, bad.ctypes.data, None, None, None) == dc.ERR_STATE
    # a disconnected seam (|dx| > 1) is not a liblqr seam with delta_x = 1 (src/render.c:313): refused, nothing carved
    jump = np.array([0, 1, 3, 3], np.int32)
    assert L.dctc_carve_and_update(c2.handle, jump.ctypes.data, None, None, None) == dc.ERR_INVALID
    assert c2.carver_size() == (6, 4)
    c2.close()


@pytest.mark.parametrize("b2,wts2", [(16, (0.5, 0.5)), (8, (0.8, 0.2)), (2, (0.5, 0.5))])
def test_set_params_rebuilds_the_map_of_a_loaded_session(ctx, b2, wts2):
    """lqr_carver_set_energy_function invalidates liblqr's energy map; here a parameter change with a session loaded
    rebuilds the resident map, so band radius / stride and map always belong to the same operator."""
    img = ol.synth_image(90, 60, 3, 41, 0)
    ctx.set_params(8, 0.5, 0.5)
    ctx.carver_load(img)
    ctx.set_params(b2, *wts2)
    assert np.array_equal(ctx.carver_energy(), ctx.energy_full(img))
    seam = np.full(60, 33, np.int32)
    ctx.carve_and_update(seam)
    assert np.array_equal(ctx.carver_energy(), ctx.energy_full(carve_host(img, seam)))
    ctx.set_params(8, 0.5, 0.5)


def test_per_pixel_symbol_serves_host_mirror(ctx):
    import ctypes as C
    img = ol.synth_image(40, 30, 3, 3, 0)
    ctx.set_params(8, 0.5, 0.5)
    ctx.carver_load(img)
    en = ctx.carver_energy()
    p = dc.CarverEnergyParams()
    p.base.edges, p.base.textures, p.base.blocksize = 0.5, 0.5, 8
    p.gpu = ctx.handle.value
    L = dc.lib()
    assert L.dctc_pixel_energy(7, 9, 40, 30, None, C.byref(p)) == en[9, 7]
    assert np.isnan(L.dctc_pixel_energy(7, 9, 41, 30, None, C.byref(p)))   # stale size: loud, not silently wrong
    ctx.carve_and_update(np.full(30, 5, np.int32), want_band=False)
    en2 = ctx.carver_energy()
    assert L.dctc_pixel_energy(7, 9, 39, 30, None, C.byref(p)) == en2[9, 7]


# ---- config 3: retarget through the host carver (liblqr stand-in) with GPU energy ---------------------------

def _naive(img, b, e, t, n, energy):
    from test_carver_cpu import naive_seams
    return naive_seams(img, b, e, t, n, energy=energy)


@pytest.mark.parametrize("device_loop", [False, True])
@pytest.mark.parametrize("b,wts", [(8, (0.5, 0.5)), (8, (0.8, 0.2)), (4, (0.5, 0.5)), (16, (0.5, 0.5))])
def test_gpu_retarget_equals_naive_loop_fed_with_gpu_energy(ctx, b, wts, device_loop):
    """Incremental GPU energy (K2) + incremental cumulative map must give exactly the seams of the naive loop
    that recomputes the full GPU energy map and the full DP for every seam."""
    from dct_carver_b200 import host
    img = ol.synth_image(150, 90, 3, 321 + b, 0)
    ctx.set_params(b, *wts)
    got = host.render(img, -25, b, *wts, ctx=ctx, device_loop=device_loop)
    seams, out = _naive(img, b, *wts, 25, energy=lambda cur: ctx.energy_full(cur))
    assert np.array_equal(got["seams"], seams)
    assert np.array_equal(got["image"], out)
    assert got["image"].shape == (90, 125, 3)


def test_gpu_retarget_seams_vs_reference_energy(ctx):
    """Seams chosen from the GPU (FP32) energies vs from the reference's double-precision energies: bit-exact until
    a near-tie in the cumulative map flips (north_star: 'bit-exact wherever energy ties do not flip').  The first
    diverging seam must BE such a tie: both loops hold the same image at that point, so with delta = the largest
    difference the energy tolerance and the FP32 sums can make along one seam, the seam the GPU loop removed costs at
    most 2*delta more than the reference's seam when both are priced with the reference's energies."""
    from dct_carver_b200 import host
    img = ol.synth_image(160, 100, 3, 77, 0)
    h = img.shape[0]
    ctx.set_params(8, 0.5, 0.5)
    got = host.render(img, -30, 8, 0.5, 0.5, ctx=ctx)
    seams_ref, _ = _naive(img, 8, 0.5, 0.5, 30, energy=lambda cur: ol.best_energy(cur, 8, 0.5, 0.5))
    same = [bool(np.array_equal(a, b)) for a, b in zip(got["seams"], seams_ref)]
    lead = same.index(False) if False in same else len(same)
    print("identical leading seams: %d / %d" % (lead, len(same)))
    assert lead >= 1
    if lead < len(same):
        cur = img
        for k in range(lead):
            cur = carve_host(cur, seams_ref[k])
        en = ol.best_energy(cur, 8, 0.5, 0.5).astype(np.float64)
        rows = np.arange(h)
        cost_gpu = en[rows, got["seams"][lead]].sum()
        cost_ref = en[rows, seams_ref[lead]].sum()
        delta = h * (ol.ABS_TOL + ol.REL_TOL * en.max()) + h * 2.0 ** -23 * max(cost_gpu, cost_ref)
        print("first divergence at seam %d: cost %.9g (GPU seam) vs %.9g (reference seam), bound %.3g" % (lead, cost_gpu, cost_ref, 2 * delta))
        assert cost_ref <= cost_gpu + 2 * delta      # the reference's seam is (nearly) optimal for its own energies
        assert cost_gpu <= cost_ref + 2 * delta      # ... and the GPU's seam ties with it within the tolerance


def test_gpu_retarget_height_and_energy_image(ctx):
    from dct_carver_b200 import host
    img = ol.synth_image(64, 70, 3, 9, 3)
    ctx.set_params(8, 0.5, 0.5)
    got = host.render(img, -6, 8, vertically=True, ctx=ctx, output_energy=True)
    assert got["image"].shape == (64, 64, 3)
    en = ctx.energy_full(img).astype(np.float32)
    with np.errstate(divide="ignore"):
        e = (np.float32(1) / (np.float32(1) + np.float32(1) / en)).astype(np.float32)
    want = (((e - e.min()) / (e.max() - e.min())) * np.float32(255)).astype(np.uint8)     # truncation, FP32 throughout
    assert np.array_equal(got["energy_image"], want)


@pytest.mark.parametrize("b,ch,w,h,n", [(8, 3, 150, 90, 25), (8, 1, 97, 75, 40), (4, 3, 64, 33, 10), (16, 3, 130, 70, 12),
                                        (8, 3, 1100, 40, 30), (8, 3, 40, 300, 8), (2, 1, 9, 1, 3),
                                        (8, 1, 4200, 20, 3), (8, 1, 5000, 40, 2), (4, 1, 13000, 12, 2), (8, 3, 3840, 70, 3), (8, 1, 2600, 50, 3)])
def test_device_seam_loop_equals_host_carver(ctx, b, ch, w, h, n):
    """dctc_carver_resize_width (seam DP + back-track + carve + band update, all on the device) must remove exactly
    the seams the host carver (liblqr stand-in, incremental cumulative map on the CPU) removes: same tie rules
    (first strict minimum among the parents, leftmost minimum of the last row), same FP32 sums.  The wide cases cover
    the 256- and 512-column strips of the seam DP, with and without bulk-copy staging of the energies."""
    from dct_carver_b200 import host
    img = ol.synth_image(w, h, ch, 500 + w, 0)
    ctx.set_params(b, 0.5, 0.5)
    want = host.render(img, -n, b, 0.5, 0.5, ctx=ctx, device_loop=False)
    ctx.set_params(b, 0.5, 0.5)
    ctx.carver_load(img)
    seams = ctx.carver_resize_width(n)
    assert np.array_equal(seams, want["seams"])
    assert ctx.carver_size() == (w - n, h)
    got_img = ctx.carver_image()
    assert np.array_equal(got_img.reshape(want["image"].shape), want["image"])
    # the resident energy plane is the full map of the carved image
    assert np.array_equal(ctx.carver_energy(), ctx.energy_full(want["image"]))


@pytest.mark.parametrize("b,ch,w,h,n,pattern", [(8, 3, 150, 90, 25, 0), (8, 1, 97, 75, 40, 0), (4, 3, 64, 33, 10, 0),
                                                (8, 3, 1100, 40, 30, 0), (8, 3, 40, 300, 8, 0), (2, 1, 9, 1, 3, 0),
                                                (8, 3, 700, 200, 30, 3), (8, 3, 640, 120, 12, 1), (8, 3, 64, 48, 6, 2)])
def test_device_seam_loop_incremental_map_equals_rebuild(ctx, b, ch, w, h, n, pattern):
    """Optional incremental cumulative map (liblqr's update_mmap on the device, dctc_carver_set_incremental): same
    seams, image and energy plane, bit for bit, as the default loop that rebuilds the map for every seam -- including
    images full of ties (gradient / checkerboard patterns), ranges that hit the image borders and the fall-back to a
    full rebuild when a row's changed range gets too wide."""
    img = ol.synth_image(w, h, ch, 700 + w, pattern)
    ctx.set_params(b, 0.5, 0.5)
    ctx.carver_load(img)
    want = ctx.carver_resize_width(n)
    want_img, want_en = ctx.carver_image(), ctx.carver_energy()
    ctx.carver_load(img)
    ctx.carver_set_incremental(True)
    try:
        # two calls: the second one continues incrementally from the map the first one left behind
        got = np.concatenate([ctx.carver_resize_width(n // 2), ctx.carver_resize_width(n - n // 2)])
        rebuilds = ctx.carver_rebuild_count()
    finally:
        ctx.carver_set_incremental(False)
    assert np.array_equal(got, want)
    assert np.array_equal(ctx.carver_image(), want_img)
    assert np.array_equal(ctx.carver_energy(), want_en)
    assert 0 <= rebuilds < n


@pytest.mark.parametrize("flat", [True, False])
@pytest.mark.parametrize("w,h", [(203, 77), (204, 77), (144, 40), (1920, 70)])
@pytest.mark.parametrize("side", ["left", "right"])
def test_device_seam_loop_hugs_the_image_border(ctx, side, w, h, flat):
    """The outermost column has the lowest energy, so every seam runs down the image border: exercises the range
    clipping of the seam DP and of the back-track windows (+inf sentinels left of column 0 and at column w) against
    the host carver.  flat: the border columns are constant (zero energy); otherwise they carry a vertical dither whose
    amplitude grows away from the border (non-zero cumulative values, so a missing sentinel would win the comparison).
    Widths that are a multiple of 4 and >= 144 make the cumulative plane's pitch equal to the width: the window of the
    first seam then ends exactly at column w (the right sentinel sits in the window row's tail pad)."""
    from dct_carver_b200 import host
    n = 4
    img = ol.synth_image(w, h, 3, 909, 0)
    k = 12
    cols = np.arange(w - k, w) if side == "right" else np.arange(k - 1, -1, -1)
    if flat:
        # window offsets are -3..+4: 4 flat columns on the right / 5 on the left make exactly the border column flat
        img[:, cols[-(4 if side == "right" else 5):], :] = 128
    else:
        amp = np.arange(k, 0, -1)                       # 12 ... 1 towards the border
        img[:, cols, :] = (128 + (np.arange(h)[:, None] & 1) * amp[None, :])[:, :, None]
    ctx.set_params(8, 0.5, 0.5)
    want = host.render(img, -n, 8, 0.5, 0.5, ctx=ctx, device_loop=False)
    ctx.set_params(8, 0.5, 0.5)
    ctx.carver_load(img)
    seams = ctx.carver_resize_width(n)
    assert np.array_equal(seams, want["seams"])
    edge = seams[0]
    assert (edge == (w - 1 if side == "right" else 0)).all(), edge
    assert np.array_equal(ctx.carver_image().reshape(want["image"].shape), want["image"])


def test_device_seam_loop_ties_and_flat_image(ctx):
    """Constant image: every energy is 0, every cumulative value ties; liblqr's rule then removes column 0 in every
    row, every time (leftmost minimum, first strict minimum among parents)."""
    img = np.full((20, 30, 3), 77, np.uint8)
    ctx.set_params(8, 0.5, 0.5)
    ctx.carver_load(img)
    seams = ctx.carver_resize_width(5)
    assert (seams == 0).all()
    assert ctx.carver_size() == (25, 20)


def test_device_seam_loop_state_errors(ctx):
    img = ol.synth_image(16, 8, 3, 3, 0)
    ctx.set_params(8, 0.5, 0.5)
    ctx.carver_load(img)
    with pytest.raises(dc.DctcError) as e:
        ctx.carver_resize_width(16)
    assert e.value.status == dc.ERR_STATE
    assert ctx.carver_resize_width(0).shape == (0, 8)


@pytest.mark.parametrize("ch,w,h", [(3, 130, 77), (1, 64, 33), (3, 1, 1)])
def test_energy_image_export_equals_host_formula(ctx, ch, w, h):
    """K3 (lqr_carver_get_energy_image semantics, src/render.c:191): 1/(1+1/e), min-max, 8-bit by truncation
    [liblqr, from memory: parity unpinned] — byte-identical to the host formula evaluated in FP32 on the same energies,
    also when (lo, hi) are supplied from outside (band sharding)."""
    img = ol.synth_image(w, h, ch, 31, 3)
    ctx.set_params(8, 0.5, 0.5)
    ctx.carver_load(img)
    en = ctx.carver_energy()
    with np.errstate(divide="ignore"):
        c = (np.float32(1.0) / (np.float32(1.0) + np.float32(1.0) / en)).astype(np.float32)
    lo, hi = c.min(), c.max()
    if hi > lo:
        want = (((c - lo) / (hi - lo)) * np.float32(255.0)).astype(np.uint8)
    else:
        want = np.zeros((h, w), np.uint8)
    got = ctx.carver_energy_image()
    assert np.array_equal(got, want)
    # two-pass form used by the row-band sharding: min/max first, then scale with the (all-reduced) pair
    d_en = ctx.dev_alloc(en.nbytes)
    d_out = ctx.dev_alloc(w * h)
    ctx.h2d(d_en, en)
    lo_hi = ctx.energy_minmax_dev(d_en, w, w, h)
    assert lo_hi[0] == lo and lo_hi[1] == hi
    ctx.energy_image_dev(d_en, w, w, h, d_out, w, lo_hi=lo_hi)
    out = np.empty((h, w), np.uint8)
    ctx.d2h(out, d_out)
    assert np.array_equal(out, want)
    ctx.dev_free(d_en)
    ctx.dev_free(d_out)


PREVIEW_GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "preview_golden.npz"))


@pytest.mark.parametrize("i", range(sum(1 for k in PREVIEW_GOLD.files if k.startswith("img_"))))
def test_preview_path_against_reference_golden(ctx, i):
    """dctc_preview_energy vs the reference's own dct_energy_preview_rows / normalize_image outputs
    (tests/golden/preview_golden.npz, generated from oracle/_ref by tests/golden/make_golden_preview.py).
    Energies live on the 0..255 luminance scale: |err| <= 255e-6 + 4e-6*|ref|; with edges != textures a near-tie
    between the classes may flip (same rule as the carver path); the 8-bit image may differ by one level where the
    FP32 energy sits on a rounding boundary."""
    img = PREVIEW_GOLD["img_%02d" % i]
    want = PREVIEW_GOLD["en_%02d" % i].astype(np.float64)
    want_img = PREVIEW_GOLD["out_%02d" % i]
    b = int(PREVIEW_GOLD["meta_%02d" % i][0])
    e, t = (float(v) for v in PREVIEW_GOLD["wts_%02d" % i])
    ctx.set_params(b, e, t)
    en, out = ctx.preview_energy(img)
    ok = np.abs(en.astype(np.float64) - want) <= 255e-6 + 4e-6 * np.abs(want)
    if e == t:
        assert ok.all(), (np.abs(en - want).max(), (~ok).sum())
    else:
        assert ok.mean() >= 0.995, (~ok).sum()
    assert out.shape == want_img.shape
    assert (out[..., 0:1] == out).all()
    diff = np.abs(out.astype(int) - want_img.astype(int))
    if e == t:
        assert diff.max() <= 1 and (diff > 0).mean() < 0.01
    ctx.set_params(8, 0.5, 0.5)


def test_preview_path_rejects_two_channels(ctx):
    img = np.zeros((4, 4, 2), np.uint8)
    with pytest.raises(dc.DctcError) as e:
        ctx.preview_energy(img)
    assert e.value.status == dc.ERR_INVALID     # convert_row_to_luminance: "Number of channels not 1 or 3"


@pytest.mark.parametrize("ch,w,h,n", [(3, 150, 90, 25), (1, 97, 40, 30)])
def test_device_vmap_and_seam_display_equal_host_carver(ctx, ch, w, h, n):
    """Visibility map (lqr_carver_set_dump_vmaps + lqr_vmap_get_data, src/render.c:214-219,374) recorded by the device
    seam loop == the host carver's, and display_carver_seams (src/render.c:204-240) painted on the device == the
    formula applied to that map."""
    from dct_carver_b200 import host
    img = ol.synth_image(w, h, ch, 700 + w, 0)
    ctx.set_params(8, 0.5, 0.5)
    want = host.render(img, -n, 8, 0.5, 0.5, ctx=ctx, output_seams=True, device_loop=False)
    assert want["vmap"] is not None and want["vmap_depth"] == n
    ctx.carver_load(img)
    ctx.carver_set_dump_vmaps(True)
    ctx.carver_resize_width(n)
    vmap, depth = ctx.carver_vmap(w, h)
    ctx.carver_set_dump_vmaps(False)
    assert depth == n
    assert np.array_equal(vmap, want["vmap"])
    assert ((vmap > 0).sum(axis=1) == n).all()
    painted = ctx.carver_paint_seams(img)
    ref = img.copy()
    if ref.ndim == 2:
        ref = ref[:, :, None]
    vis = vmap[:h - 1, :w - 1]
    ys, xs = np.nonzero(vis)
    ref[ys, xs, 0] = 0
    if ch > 1:
        ref[ys, xs, 1] = (255.0 * vis[ys, xs].astype(np.float64) / float(depth)).astype(np.uint8)
    if ch > 2:
        ref[ys, xs, 2] = 0
    assert np.array_equal(painted, ref)


@pytest.mark.parametrize("ch,w,h,k", [(3, 150, 90, 25), (1, 97, 40, 30), (4, 64, 33, 9), (3, 1930, 64, 12)])
def test_device_seam_enlarging_equals_host_carver(ctx, ch, w, h, k):
    """seams_number > 0 (src/render.c:357-363; lqr_carver_resize to a larger size): the device path (seam loop + inflate
    kernel) gives the image of the host carver's enlarge (host seam loop) and of the numpy restatement of
    lqr_carver_inflate applied to the device's visibility map [pixel synthesis from memory: parity unpinned]; the
    seams are those of the shrink by the same amount; the session continues on the enlarged image."""
    from dct_carver_b200 import host
    from test_carver_cpu import inflate_from_vmap
    img = ol.synth_image(w, h, ch, 1234 + w, 0)
    ctx.set_params(8, 0.5, 0.5)
    want = host.render(img, +k, 8, 0.5, 0.5, ctx=ctx, device_loop=False)
    got = host.render(img, +k, 8, 0.5, 0.5, ctx=ctx, device_loop=True)
    assert got["image"].shape == (h, w + k, ch)
    assert np.array_equal(got["seams"], want["seams"])
    assert np.array_equal(got["vmap"], want["vmap"]) and got["vmap_depth"] == k
    assert np.array_equal(got["image"], want["image"])
    assert np.array_equal(got["image"], inflate_from_vmap(img.reshape(h, w, ch), got["vmap"], k))
    # straight through the C ABI: seams of the shrink, session continues on the enlarged frame with a fresh energy map
    ctx.carver_load(img)
    shrink = ctx.carver_resize_width(k)
    ctx.carver_load(img)
    seams = ctx.carver_enlarge_width(k)
    assert np.array_equal(seams, shrink)
    assert ctx.carver_size() == (w + k, h)
    assert np.array_equal(ctx.carver_image(), got["image"])
    assert np.array_equal(ctx.carver_energy(), ctx.energy_full(got["image"]))
    vmap, depth = ctx.carver_vmap(w, h)
    assert depth == k and np.array_equal(vmap, got["vmap"])


def test_device_seam_enlarging_state_errors(ctx):
    img = ol.synth_image(16, 8, 3, 3, 0)
    ctx.set_params(8, 0.5, 0.5)
    ctx.carver_load(img)
    with pytest.raises(dc.DctcError) as e:
        ctx.carver_enlarge_width(16)              # at most w - 1 seams per pass
    assert e.value.status == dc.ERR_STATE
    ctx.carver_resize_width(2)
    with pytest.raises(dc.DctcError) as e:
        ctx.carver_enlarge_width(3)               # needs a freshly loaded frame
    assert e.value.status == dc.ERR_STATE
