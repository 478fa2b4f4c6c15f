"""CPU tests: the oracle restatement against (a) the reference's own C compiled unmodified (oracle/_ref,
when present), (b) the committed golden vectors generated from it, (c) the survey's known-answer numbers."""
import os

import numpy as np
import pytest

import oracle_lib as ol

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "energy_golden.npz"))
N_GOLD = sum(1 for k in GOLD.files if k.startswith("img_"))


def _same_to_one_ulp(got, want):
    """Equal up to one float32 ulp, with a 1e-12 floor: on flat regions the reference's butterflies give an exact 0
    where the closed form leaves O(1e-16) residue (and vice versa for the FFT-based b=2,4 path)."""
    d = np.abs(got.astype(np.float64) - want.astype(np.float64))
    return bool((d <= 1e-12 + 1.2e-7 * np.abs(want.astype(np.float64))).all())


def _exact_fraction(got, want):
    return float((got == want).mean())


@pytest.mark.parametrize("i", range(N_GOLD))
def test_oracle_matches_golden(i):
    img = GOLD["img_%02d" % i]
    want = GOLD["en_%02d" % i]
    b = int(GOLD["meta_%02d" % i][0])
    e, t = (float(v) for v in GOLD["wts_%02d" % i])
    got = ol.oracle_energy(img, b, e, t, nthreads=2)
    assert _same_to_one_ulp(got, want)
    if int(GOLD["meta_%02d" % i][1]) == 0 and b >= 4:
        assert _exact_fraction(got, want) == 1.0   # bit-exact on noise


def test_golden_inputs_regenerate():
    """The committed inputs are exactly what the synthetic generator produces (so GPU tests can regenerate them)."""
    for i in range(N_GOLD):
        b, pattern, ch, idx = (int(v) for v in GOLD["meta_%02d" % i])
        img = ol.synth_image(64, 48, ch, 0xD0C7CA12 + idx, pattern)
        assert np.array_equal(img, GOLD["img_%02d" % i])


def test_survey_known_answers():
    """LCG luma plane 512x512, b=8, e=t=.5 -> sum 99646.634, max 1.04857135, E[0]=0.682583332."""
    s = 12345
    luma = np.empty(512 * 512)
    for k in range(luma.size):
        s = (s * 1664525 + 1013904223) & 0xFFFFFFFF
        luma[k] = (s >> 24) / 255.0
    en = ol.oracle_energy_luma(luma.reshape(512, 512), 8, 0.5, 0.5)
    kat = GOLD["kat_lcg_b8_sum"]
    assert abs(en.astype(np.float64).sum() - 99646.634) < 1e-2
    assert abs(en.astype(np.float64).sum() - kat[0]) < 1e-6
    assert abs(en.max() - 1.04857135) < 1e-7
    assert abs(en.flat[0] - 0.682583332) < 1e-7 and abs(en.flat[1] - 0.553684235) < 1e-7
    assert abs(en.flat[513] - 0.546698153) < 1e-7 and abs(en.flat[-1] - 0.377450973) < 1e-7


@pytest.mark.skipif(ol.ref() is None, reason="oracle/_ref not built (reference sources absent)")
@pytest.mark.parametrize("b", [2, 4, 8, 16])
@pytest.mark.parametrize("wts", [(0.5, 0.5), (0.8, 0.2), (0.0, 1.0)])
def test_oracle_vs_compiled_reference(b, wts):
    for pattern, ch, (w, h) in ((0, 3, (83, 37)), (3, 1, (40, 70)), (1, 4, (33, 33)), (2, 3, (96, 17)), (0, 1, (5, 3)),
                                (0, 3, (1, 1)), (0, 2, (2, 9))):
        img = ol.synth_image(w, h, ch, 77 + b, pattern)
        a = ol.ref_energy(img, b, wts[0], wts[1], nthreads=1)
        o = ol.oracle_energy(img, b, wts[0], wts[1], nthreads=3)
        assert _same_to_one_ulp(o, a), (pattern, ch, w, h)


def test_basis_normalisation_quirk():
    """b=8,16 orthonormal; b=2,4 unnormalised."""
    import ctypes as C
    for b in (2, 4, 8, 16):
        B = np.zeros((b, b))
        ol.oracle().dctc_oracle_basis(b, B.ctypes.data_as(C.POINTER(C.c_double)))
        G = B @ B.T
        if b >= 8:
            assert np.allclose(G, np.eye(b), atol=1e-14)
        else:
            assert np.allclose(G, np.diag([b] + [b / 2.0] * (b - 1)), atol=1e-14)
            assert np.allclose(B[0], 1.0)


def test_constant_image_has_zero_energy():
    for b in (4, 8, 16):
        img = np.full((20, 31, 3), 173, np.uint8)
        assert ol.oracle_energy(img, b).max() < 1e-12


def test_edge_atom_classes_on_pure_steps():
    """A vertical step edge excites only horizontal frequencies k1>0,k2=0: arg-max is the edge atom (1,0)."""
    img = np.zeros((32, 32), np.uint8)
    img[:, 16:] = 255
    en, cls = ol.oracle_energy(img, 8, 0.9, 0.1, want_class=True)
    assert cls[10, 14] == 1 and en[10, 14] > 0
    assert en[10, 2] == 0.0
    en2, cls2 = ol.oracle_energy(img.T.copy(), 8, 0.9, 0.1, want_class=True)
    assert np.allclose(en2, en.T, atol=1e-7)


def test_preview_operator_window_and_luminance():
    """Preview path (render.c:31-79): window [-(C-1), b-C], BT.601 byte luminance, double output."""
    import ctypes as C
    img = ol.synth_image(24, 20, 3, 5, 0)
    lum = np.zeros((20, 24), np.uint8)
    ol.oracle().dctc_oracle_preview_luminance(img.ctypes.data_as(C.POINTER(C.c_uint8)), 24, 20, 3, C.c_size_t(72),
                                              lum.ctypes.data_as(C.POINTER(C.c_uint8)))
    want = (16.0 + img[..., 0] * 0.2568 + img[..., 1] * 0.5041 + img[..., 2] * 0.0979).astype(np.uint8)
    assert np.array_equal(lum, want)
    out = np.zeros((20, 24))
    rc = ol.oracle().dctc_oracle_preview_energy(lum.ctypes.data_as(C.POINTER(C.c_uint8)), 24, 20, 8, C.c_float(0.5),
                                                C.c_float(0.5), out.ctypes.data_as(C.POINTER(C.c_double)))
    assert rc == 0
    # independent dense evaluation at one interior pixel: window rows/cols y-2..y+5
    B = np.array([[np.sqrt(2 / 8) * (np.sqrt(.5) if k == 0 else 1) * np.cos(np.pi * (j + .5) * k / 8) for j in range(8)]
                  for k in range(8)])
    y, x = 9, 11
    D = lum[y - 2:y + 6, x - 2:x + 6].astype(np.float64)
    T = np.abs(B @ D @ B.T)
    T[0, 0] = 0
    assert abs(out[y, x] - 0.5 * T.max()) < 1e-9


@pytest.mark.parametrize("b", [2, 4, 8, 16])
@pytest.mark.parametrize("wts", [(0.5, 0.5), (0.8, 0.2)])
@pytest.mark.parametrize("ch,w,h", [(3, 37, 29), (1, 24, 20), (4, 18, 9)])
def test_preview_oracle_pinned_to_reference(b, wts, ch, w, h):
    """The oracle's preview restatement against the reference's OWN dct_energy_preview_rows /
    convert_row_to_luminance / normalize_image (src/render.c:31-109), driven like dct_energy_preview."""
    if ol.ref() is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    img = ol.synth_image(w, h, ch, 40 + b, 0)
    en_ref, img_ref = ol.ref_preview(img, b, *wts)
    en_orc, lum = ol.oracle_preview(img, b, *wts)
    # the preview map holds float results of weighted_max_dct_correlation stored in doubles
    assert np.array_equal(en_ref.astype(np.float32).astype(np.float64), en_ref)
    # the energies are O(100) on the 0..255 luminance scale: agreement to a float ulp, up to class flips on exact ties
    close = np.abs(en_orc - en_ref) <= 2e-5 * np.maximum(np.abs(en_ref), 1.0)
    assert close.mean() > 0.995, (~close).sum()
    if wts[0] == wts[1]:
        assert close.all()
    assert np.array_equal(ol.preview_normalize(en_ref, ch), img_ref)


def test_preview_golden_fixture_matches_oracle():
    """The committed preview fixture (generated from the compiled reference) against the oracle restatement, so the
    check also runs where /root/reference is absent."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "preview_golden.npz"))
    n = sum(1 for k in g.files if k.startswith("img_"))
    assert n >= 8
    for i in range(n):
        img = g["img_%02d" % i]
        b = int(g["meta_%02d" % i][0])
        e, t = (float(v) for v in g["wts_%02d" % i])
        en, _ = ol.oracle_preview(img, b, e, t)
        want = g["en_%02d" % i].astype(np.float64)
        close = np.abs(en - want) <= 2e-5 * np.maximum(np.abs(want), 1.0)
        assert close.mean() > 0.995 and (e != t or close.all()), i
        assert np.array_equal(ol.preview_normalize(want, img.shape[2] if img.ndim == 3 else 1), g["out_%02d" % i])
