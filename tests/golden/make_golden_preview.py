#!/usr/bin/env python3
"""Generates tests/golden/preview_golden.npz from the reference's OWN preview path (oracle/_ref, i.e.
/root/reference/src/render.c:31-109 compiled where it lies; see oracle/Makefile): for each case the input image,
the un-normalised energy map (float32, exactly what the reference's gdouble plane holds) and the normalised image.
Run in the build container (needs /root/reference); the fixture is what travels to the GPU box."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol  # noqa: E402

CASES = [  # (pattern, channels, w, h, blocksize, edges, textures)
    (0, 3, 61, 45, 8, 0.5, 0.5), (0, 1, 40, 33, 8, 0.5, 0.5), (3, 3, 70, 38, 8, 0.8, 0.2), (0, 3, 33, 21, 2, 0.5, 0.5),
    (0, 3, 50, 30, 4, 0.5, 0.5), (0, 3, 47, 41, 16, 0.5, 0.5), (1, 4, 36, 20, 8, 0.3, 0.7), (0, 3, 5, 3, 8, 0.5, 0.5),
]


def main():
    assert ol.ref() is not None, "oracle/_ref missing: run `make -C oracle` where /root/reference exists"
    out = {}
    for i, (pat, ch, w, h, b, e, t) in enumerate(CASES):
        img = ol.synth_image(w, h, ch, 900 + i, pat)
        en, im = ol.ref_preview(img, b, e, t)
        assert np.array_equal(en.astype(np.float32).astype(np.float64), en)
        out["img_%02d" % i] = img
        out["en_%02d" % i] = en.astype(np.float32)
        out["out_%02d" % i] = im
        out["meta_%02d" % i] = np.array([b], np.int32)
        out["wts_%02d" % i] = np.array([e, t], np.float32)
    np.savez_compressed(os.path.join(HERE, "preview_golden.npz"), **out)
    print("wrote", len(CASES), "cases")


if __name__ == "__main__":
    main()
