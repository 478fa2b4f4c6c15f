"""CPU tests of the host-side carver (the liblqr stand-in) driven by a per-pixel callback — the reference's own
dct_pixel_energy from oracle/_ref — against the independent naive restatement in oracle/oracle_carver.c."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as ol
from dct_carver_b200 import host

pytestmark = pytest.mark.skipif(ol.ref() is None, reason="oracle/_ref (compiled reference) not present")


class RefEnergyParameters(C.Structure):   # src/render.h:9-18
    _fields_ = [("edges", C.c_float), ("textures", C.c_float), ("blocksize", C.c_int), ("ip", C.c_void_p),
                ("w", C.c_void_p), ("data", C.c_void_p)]


def ref_callback(b, e, t):
    """The reference's callback + its scratch exactly as src/render.c:296-305 allocates it."""
    R = ol.ref()
    for f in ("alloc_1d_int", "alloc_1d_double", "alloc_2d_double"):
        getattr(R, f).restype = C.c_void_p
    ep = RefEnergyParameters(e, t, b, R.alloc_1d_int(2 + int(np.sqrt(b / 2 + 0.5))), R.alloc_1d_double(b * 3 // 2),
                             R.alloc_2d_double(b, b))
    C.cast(ep.ip, C.POINTER(C.c_int))[0] = 0
    return C.cast(R.dct_pixel_energy, C.c_void_p), ep


def naive_seams(img, b, e, t, n, energy=None):
    img = np.ascontiguousarray(img)
    h, w, ch = img.shape
    seams = np.zeros((n, h), np.int32)
    out = np.zeros((h, w - n, ch), np.uint8)
    FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p)

    def cb(p, cw, chh, cch, outp, user):
        cur = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), (chh, cw, cch))
        en = energy(cur.copy())
        C.memmove(outp, en.ctypes.data, en.nbytes)
        return 0
    fn = FN(cb) if energy is not None else None
    rc = ol.oracle().dctc_oracle_retarget_width(img.ctypes.data_as(C.c_void_p), w, h, ch, b, C.c_float(e), C.c_float(t), n,
                                                seams.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), fn, None, 2)
    assert rc == 0
    return seams, out


@pytest.mark.parametrize("b,wts,ch", [(8, (0.5, 0.5), 3), (8, (0.8, 0.2), 1), (4, (0.5, 0.5), 3), (16, (0.5, 0.5), 1)])
def test_incremental_carver_matches_naive_restatement(b, wts, ch):
    img = ol.synth_image(57, 41, ch, 900 + b, 0)
    cbp, ep = ref_callback(b, *wts)
    got = host.render(img, -9, b, *wts, callback=cbp, callback_extra=C.byref(ep))
    seams, out = naive_seams(img, b, *wts, 9, energy=lambda cur: ol.ref_energy(cur, b, *wts, nthreads=1))
    assert np.array_equal(got["seams"], seams)
    assert np.array_equal(got["image"], out)
    assert got["vmap_depth"] == 9 and (got["vmap"] > 0).sum() == 9 * 41
    # every seam is connected (delta_x = 1) and inside the image at removal time
    for k, s in enumerate(got["seams"]):
        assert np.abs(np.diff(s)).max() <= 1 and s.min() >= 0 and s.max() < 57 - k


def test_height_retarget_is_width_retarget_of_the_transpose():
    img = ol.synth_image(33, 45, 3, 5, 0)
    cbp, ep = ref_callback(8, 0.5, 0.5)
    a = host.render(img, -5, 8, vertically=True, callback=cbp, callback_extra=C.byref(ep))
    bres = host.render(np.ascontiguousarray(img.transpose(1, 0, 2)), -5, 8, vertically=False, callback=cbp,
                       callback_extra=C.byref(ep))
    assert a["image"].shape == (40, 33, 3)
    assert np.array_equal(a["image"], bres["image"].transpose(1, 0, 2))


def test_render_refuses_without_gpu_or_callback():
    import dct_carver_b200 as dc
    img = ol.synth_image(20, 20, 3, 1, 0)
    with pytest.raises(dc.DctcError) as e:
        host.render(img, -2)
    assert e.value.status == dc.ERR_NO_DEVICE        # no CPU fallback


def inflate_from_vmap(img, vmap, k):
    """liblqr's lqr_carver_inflate restated in numpy [from memory: parity unpinned]: every pixel of the first k seams is
    doubled, the new pixel on its left = integer mean of the pixel and its left neighbour (a copy in column 0)."""
    h, w, ch = img.shape
    out = np.zeros((h, w + k, ch), np.uint8)
    for y in range(h):
        n = 0
        for x in range(w):
            if 0 < vmap[y, x] <= k:
                out[y, x + n] = ((img[y, x - 1].astype(int) + img[y, x].astype(int)) // 2) if x > 0 else img[y, x]
                n += 1
            out[y, x + n] = img[y, x]
        assert n == k
    return out


@pytest.mark.parametrize("vertically", [False, True])
def test_enlarging_duplicates_the_seams_a_shrink_would_remove(vertically):
    """seams_number > 0 (src/render.c:357-363): same seams, same order as the shrink by the same amount; pixel synthesis
    by duplicate-and-average."""
    img = ol.synth_image(41, 37, 3, 77, 0)
    k = 7
    cbp, ep = ref_callback(8, 0.5, 0.5)
    grow = host.render(img, +k, 8, vertically=vertically, callback=cbp, callback_extra=C.byref(ep))
    shrink = host.render(img, -k, 8, vertically=vertically, callback=cbp, callback_extra=C.byref(ep))
    assert np.array_equal(grow["seams"], shrink["seams"])
    assert np.array_equal(grow["vmap"], shrink["vmap"]) and grow["vmap_depth"] == k
    if vertically:
        want = inflate_from_vmap(np.ascontiguousarray(img.transpose(1, 0, 2)), grow["vmap"], k).transpose(1, 0, 2)
        assert grow["image"].shape == (37 + k, 41, 3)
    else:
        want = inflate_from_vmap(img, grow["vmap"], k)
        assert grow["image"].shape == (37, 41 + k, 3)
    assert np.array_equal(grow["image"], want)


def test_enlarging_beyond_twice_the_width_runs_in_passes():
    img = ol.synth_image(12, 9, 1, 5, 0)
    cbp, ep = ref_callback(4, 0.5, 0.5)
    got = host.render(img, +15, 4, callback=cbp, callback_extra=C.byref(ep))    # 12 -> 23 (+11), then 23 -> 27 (+4)
    assert got["image"].shape == (9, 27, 1)
