"""CPU tests of the boundary: the C-ABI library loads, exports every symbol include/dctc.h declares, keeps the
reference's EnergyParameters layout, and fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import dct_carver_b200 as dc
import oracle_lib as ol

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "dctc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dctc_[a-z0-9_]+)\s*\(", src)))


def test_header_and_python_mirror_agree():
    assert _header_symbols() == sorted(dc.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol():
    L = C.CDLL(dc.LIB_PATH)
    for name in _header_symbols():
        assert hasattr(L, name), name
    assert dc.lib().dctc_version() == 1


def test_energy_parameters_layout_matches_reference_struct():
    """src/render.h:9-18: gfloat edges; gfloat textures; gint blocksize; int* ip; double* w; double** data."""
    E = dc.EnergyParameters
    assert (E.edges.offset, E.textures.offset, E.blocksize.offset) == (0, 4, 8)
    assert (E.ip.offset, E.w.offset, E.data.offset) == (16, 24, 32)
    assert C.sizeof(E) == 40
    assert dc.CarverEnergyParams.gpu.offset == 40


def test_strerror_and_argument_errors_without_device():
    L = dc.lib()
    assert L.dctc_strerror(0) == b"ok"
    assert b"blocksize" in L.dctc_strerror(dc.ERR_BLOCKSIZE)
    assert b"no CPU fallback" in L.dctc_strerror(dc.ERR_NO_DEVICE)
    assert L.dctc_create(None, 0) == dc.ERR_INVALID
    # NULL context: every compute entry refuses instead of computing on the CPU
    out = np.zeros(4, np.float32)
    img = np.zeros(4, np.uint8)
    assert L.dctc_energy_full(None, img.ctypes.data, 2, 2, 1, 2, out.ctypes.data) == dc.ERR_INVALID
    assert L.dctc_carve_and_update(None, None, None, None, None) == dc.ERR_INVALID
    assert np.isnan(L.dctc_pixel_energy(0, 0, 2, 2, None, None))


def test_no_cpu_fallback_when_no_device():
    L = dc.lib()
    if L.dctc_device_count() > 0:
        pytest.skip("a CUDA device is present")
    h = C.c_void_p()
    assert L.dctc_create(C.byref(h), 0) == dc.ERR_NO_DEVICE
    assert not h.value
    with pytest.raises(dc.DctcError):
        dc.Context()


def test_synth_generator_host_matches_numpy_mirror():
    L = dc.lib()
    for pattern in range(4):
        img = ol.synth_image(70, 9, 3, 0xABCDEF01, pattern, frame=3, y_offset=60)
        for (y, x, c) in ((0, 0, 0), (8, 69, 2), (4, 33, 1), (7, 64, 0), (3, 7, 2)):
            assert L.dctc_synth_byte(0xABCDEF01, 3, y + 60, x, c, pattern) == img[y, x, c], (pattern, y, x, c)


def test_product_does_not_reference_the_oracle():
    """The product path must never import / link / call anything under oracle/."""
    pkg = os.path.join(ROOT, "dct_carver_b200")
    for base, _, files in os.walk(pkg):
        if os.path.basename(base) == "build":
            continue
        for f in files:
            if f.endswith((".cu", ".cuh", ".h", ".c", ".py")):
                text = open(os.path.join(base, f), errors="ignore").read()
                assert "liboracle" not in text and "libdctc_ref" not in text and "oracle_dct" not in text, f
                assert not re.search(r"(import|from)\s+oracle", text), f
