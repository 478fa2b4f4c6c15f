import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))
import numpy as np
import dct_carver_b200 as dc, oracle_lib as ol
ctx = dc.Context(0)
for (pattern, ch, w, h) in [(0, 1, 208, 77), (0, 1, 16, 3), (0, 1, 1008, 40), (0,1,256,16), (0,1,512,16)]:
    for b in (2, 4):
        img = ol.synth_image(w, h, ch, 2000 + b, pattern)
        ctx.set_params(b, 0.5, 0.5)
        ctx.set_kernel(dc.KERNEL_AUTO); got = ctx.energy_full(img)
        ctx.set_kernel(dc.KERNEL_FP32_TILE); tile = ctx.energy_full(img)
        ctx.set_kernel(dc.KERNEL_AUTO)
        d = np.argwhere(got.view(np.uint32) != tile.view(np.uint32))
        print(b, (pattern, ch, w, h), "ndiff", len(d), "first", d[:6].tolist(), "cols", sorted(set(d[:,1].tolist()))[:20] if len(d) else [])
        if len(d):
            y, x = d[0]; print("   got %r tile %r" % (got[y, x], tile[y, x]), "maxabs", np.abs(got-tile).max())
