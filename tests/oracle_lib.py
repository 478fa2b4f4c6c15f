"""Test-side access to the checker: oracle/liboracle.so (our CPU restatement) and, when it was built,
oracle/_ref/libdctc_ref.so (the reference's own C hot path compiled unmodified).  Also a numpy mirror of the
synthetic image generator (dctc_synth_px) so CPU and GPU see identical inputs."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_ORC = None
_REF = None

_dp = C.POINTER(C.c_double)
_fp = C.POINTER(C.c_float)
_bp = C.POINTER(C.c_uint8)


def oracle():
    global _ORC
    if _ORC is None:
        _ORC = C.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
    return _ORC


def ref():
    """The compiled reference, or None when oracle/_ref was not built (e.g. /root/reference absent)."""
    global _REF
    if _REF is None:
        p = os.path.join(ROOT, "oracle", "_ref", "libdctc_ref.so")
        if not os.path.exists(p):
            return None
        _REF = C.CDLL(p)
    return _REF


def _img3(img):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    if img.ndim == 2:
        img = img[:, :, None]
    return img


def oracle_energy(img, b=8, edges=0.5, textures=0.5, nthreads=8, want_class=False):
    img = _img3(img)
    h, w, ch = img.shape
    out = np.zeros((h, w), np.float32)
    cls = np.zeros((h, w), np.uint8) if want_class else None
    rc = oracle().dctc_oracle_energy_image(img.ctypes.data_as(_bp), w, h, ch, C.c_size_t(w * ch), b, C.c_float(edges),
                                           C.c_float(textures), out.ctypes.data_as(_fp),
                                           cls.ctypes.data_as(_bp) if want_class else None, nthreads)
    assert rc == 0, rc
    return (out, cls) if want_class else out


def oracle_energy_luma(luma, b=8, edges=0.5, textures=0.5, nthreads=8):
    luma = np.ascontiguousarray(luma, dtype=np.float64)
    h, w = luma.shape
    out = np.zeros((h, w), np.float32)
    rc = oracle().dctc_oracle_energy_rows(luma.ctypes.data_as(_dp), w, h, b, C.c_float(edges), C.c_float(textures),
                                          out.ctypes.data_as(_fp), None, 0, h, nthreads)
    assert rc == 0, rc
    return out


def ref_energy(img, b=8, edges=0.5, textures=0.5, nthreads=8):
    img = _img3(img)
    h, w, ch = img.shape
    out = np.zeros((h, w), np.float32)
    rc = ref().dctc_ref_energy_image(img.ctypes.data_as(_bp), w, h, ch, C.c_size_t(w * ch), b, C.c_float(edges),
                                     C.c_float(textures), out.ctypes.data_as(_fp), nthreads)
    assert rc == 0, rc
    return out


def ref_energy_luma(luma, b=8, edges=0.5, textures=0.5, nthreads=8):
    luma = np.ascontiguousarray(luma, dtype=np.float64)
    h, w = luma.shape
    out = np.zeros((h, w), np.float32)
    ref().dctc_ref_energy_rows(luma.ctypes.data_as(_dp), w, h, b, C.c_float(edges), C.c_float(textures),
                               out.ctypes.data_as(_fp), 0, h, nthreads)
    return out


def best_energy(img, b=8, edges=0.5, textures=0.5):
    """The strongest checker available: the compiled reference if present, else the restatement."""
    if ref() is not None:
        return ref_energy(img, b, edges, textures)
    return oracle_energy(img, b, edges, textures)


# ---- numpy mirror of dctc_synth_px (dct_carver_b200/csrc/dctc_common.cuh) ---------------------------------

def _mix32(h):
    h = h.astype(np.uint64)
    h ^= h >> np.uint64(16)
    h = (h * np.uint64(0x85EBCA6B)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(13)
    h = (h * np.uint64(0xC2B2AE35)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(16)
    return h


def synth_image(w, h, ch, seed, pattern=0, frame=0, y_offset=0):
    M = np.uint64(0xFFFFFFFF)
    y = (np.arange(h, dtype=np.uint64) + np.uint64(y_offset))[:, None, None]
    x = np.arange(w, dtype=np.uint64)[None, :, None]
    c = np.arange(ch, dtype=np.uint64)[None, None, :]
    frame = np.uint64(frame)
    seed = np.uint64(seed)
    if pattern == 0:
        hh = (seed ^ ((frame * np.uint64(0x9E3779B1)) & M)) & M
        hh = _mix32(hh ^ ((y * np.uint64(0x85EBCA77)) & M))
        hh = _mix32(hh ^ ((x * np.uint64(0xC2B2AE3D)) & M))
        hh = _mix32(hh ^ ((c * np.uint64(0x27D4EB2F)) & M))
        return (hh >> np.uint64(24)).astype(np.uint8)
    if pattern == 1:
        t = (x + np.uint64(2) * y + np.uint64(5) * frame + np.uint64(3) * c) & np.uint64(511)
        tri = np.where(t < 256, t, np.uint64(511) - t)
        g = ((x >> np.uint64(3)) + (y >> np.uint64(4)) + (seed & np.uint64(15))) & np.uint64(255)
        return ((tri + g + np.uint64(0) * c) >> np.uint64(1)).astype(np.uint8)
    if pattern == 2:
        cell = ((x >> np.uint64(3)) + (y >> np.uint64(3)) + frame) & np.uint64(1)
        flat = (((x >> np.uint64(6)) + (y >> np.uint64(6))) % np.uint64(3)) == 0
        v = np.where(cell == 1, np.uint64(200) + np.uint64(10) * c, np.uint64(40) + np.uint64(0) * c)
        return np.where(flat, np.uint64(128), v).astype(np.uint8)
    if pattern == 3:
        xi = (x & np.uint64(63)).astype(np.int64) - 32
        yi = (y & np.uint64(63)).astype(np.int64) - 32
        o = (((x >> np.uint64(6)) + np.uint64(3) * (y >> np.uint64(6)) + frame) & np.uint64(7)).astype(np.int64)
        s = np.select([o == 0, o == 1, o == 2, o == 3, o == 4, o == 5, o == 6],
                      [xi + 0 * yi, yi + 0 * xi, xi + yi, xi - yi, 2 * xi + yi, xi - 2 * yi, 3 * xi + yi], xi + 3 * yi)
        ci = c.astype(np.int64)
        return np.where(s >= 0, 220 - 20 * ci, 30 + 5 * ci).astype(np.uint8)
    raise ValueError(pattern)


def parity(got, want):
    """max abs error, max rel error with the 1e-6 denominator floor of SURVEY section 8(d)."""
    got = got.astype(np.float64)
    want = want.astype(np.float64)
    d = np.abs(got - want)
    return float(d.max()), float((d / np.maximum(np.abs(want), 1e-6)).max())


# Stated tolerance of the FP32 CUDA path against the double-precision reference:
# |err| <= ABS_TOL + REL_TOL*|ref|.  Energies are O(0.01..2); FP32 rounding through two 1-D transforms of
# 0..255-unit luma gives ~1e-7 relative to the DC scale, hence the absolute floor.
REL_TOL = 4e-6
ABS_TOL = 1e-6


def assert_parity(got, want, rel=REL_TOL, abs_=ABS_TOL):
    got = got.astype(np.float64)
    want = want.astype(np.float64)
    d = np.abs(got - want)
    bad = d > abs_ + rel * np.abs(want)
    assert not bad.any(), "parity: %d px off, max abs %.3e at %s (got %r want %r)" % (
        bad.sum(), d.max(), np.unravel_index(d.argmax(), d.shape), got.flat[d.argmax()], want.flat[d.argmax()])


def ref_preview(img, b=8, edges=0.5, textures=0.5, want_image=True):
    """The reference's own preview path (dct_energy_preview_rows + convert_row_to_luminance + normalize_image,
    src/render.c:31-109) driven like dct_energy_preview: returns (energy float64 map, normalised image or None)."""
    L = ref()
    if L is None:
        return None
    img = _img3(img)
    h, w, ch = img.shape
    en = np.zeros((h, w), np.float64)
    out = np.zeros((h, w, ch), np.uint8) if want_image else None
    L.dctc_ref_preview.restype = C.c_int
    rc = L.dctc_ref_preview(img.ctypes.data_as(C.c_void_p), w, h, ch, C.c_size_t(w * ch), b, C.c_float(edges),
                            C.c_float(textures), en.ctypes.data_as(C.c_void_p),
                            out.ctypes.data_as(C.c_void_p) if want_image else None)
    assert rc == 0
    return en, out


def oracle_preview(img, b=8, edges=0.5, textures=0.5):
    """Oracle restatement of the preview operator: (energy float64 map, luminance bytes)."""
    L = oracle()
    img = _img3(img)
    h, w, ch = img.shape
    lum = np.zeros((h, w), np.uint8)
    L.dctc_oracle_preview_luminance(img.ctypes.data_as(C.c_void_p), w, h, ch, C.c_size_t(w * ch), lum.ctypes.data_as(C.c_void_p))
    en = np.zeros((h, w), np.float64)
    rc = L.dctc_oracle_preview_energy(lum.ctypes.data_as(C.c_void_p), w, h, b, C.c_float(edges), C.c_float(textures),
                                      en.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return en, lum


def preview_normalize(en, channels):
    """normalize_image (src/render.c:81-109) in float64: (guchar) ROUND(255 * ((e - min) / (max - min)))."""
    en = np.asarray(en, np.float64)
    lo, hi = en.min(), en.max()
    if hi > lo:
        v = (255.0 * ((en - lo) / (hi - lo)) + 0.5).astype(np.int64).astype(np.uint8)
    else:
        v = np.zeros(en.shape, np.uint8)
    return np.repeat(v[:, :, None], channels, axis=2)
