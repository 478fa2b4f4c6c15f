#!/usr/bin/env python3
"""Generates tests/golden/energy_golden.npz from the REFERENCE ITSELF: oracle/_ref/libdctc_ref.so is the
reference's unmodified src/dct.c + src/render.c + src/fft2d/*.c compiled by oracle/Makefile.  Run in the
authoring container (needs /root/reference); the .npz travels with the repo, /root/reference does not.

    make -C oracle && python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol  # noqa: E402

W, H = 64, 48
SEED = 0xD0C7CA12


def cases():
    out = []
    for b in (2, 4, 8, 16):
        for (e, t) in ((0.5, 0.5), (0.8, 0.2)):
            out.append(dict(b=b, e=e, t=t, pattern=0, ch=3))
            out.append(dict(b=b, e=e, t=t, pattern=3, ch=1))
    for pattern, ch in ((1, 3), (2, 3), (0, 4), (0, 2), (0, 1)):
        out.append(dict(b=8, e=0.5, t=0.5, pattern=pattern, ch=ch))
    return out


def main():
    assert ol.ref() is not None, "build oracle/_ref first (make -C oracle)"
    store = {}
    for i, c in enumerate(cases()):
        img = ol.synth_image(W, H, c["ch"], SEED + i, c["pattern"])
        en = ol.ref_energy(img, c["b"], c["e"], c["t"], nthreads=1)
        store["img_%02d" % i] = img
        store["en_%02d" % i] = en
        store["meta_%02d" % i] = np.array([c["b"], c["pattern"], c["ch"], i], np.int32)
        store["wts_%02d" % i] = np.array([c["e"], c["t"]], np.float32)
    # known-answer vector on a double luma plane (SURVEY section 8c LCG): 512x512, b=8, e=t=.5
    s = 12345
    luma = np.empty(512 * 512)
    for k in range(luma.size):
        s = (s * 1664525 + 1013904223) & 0xFFFFFFFF
        luma[k] = (s >> 24) / 255.0
    en = ol.ref_energy_luma(luma.reshape(512, 512), 8, 0.5, 0.5)
    store["kat_lcg_b8_sum"] = np.array([en.astype(np.float64).sum(), en.max(), en.flat[0], en.flat[1], en.flat[513], en.flat[-1]])
    np.savez_compressed(os.path.join(HERE, "energy_golden.npz"), **store)
    print("wrote", len(cases()), "cases;", os.path.getsize(os.path.join(HERE, "energy_golden.npz")), "bytes")


if __name__ == "__main__":
    main()
