import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
import os
import dct_carver_b200 as dc, oracle_lib as ol
if os.environ.get("DCTC_LIB"): dc.LIB_PATH = os.environ["DCTC_LIB"]
from dct_carver_b200 import host
ctx = dc.Context(0, kernel=dc.KERNEL_FP32_MARCH)
for (b, ch, w, h, n) in [(8, 3, 1100, 40, 30)]:
    img = ol.synth_image(w, h, ch, 500 + w, 0)
    ctx.set_params(b, 0.5, 0.5)
    want = host.render(img, -n, b, 0.5, 0.5, ctx=ctx)
    ctx.carver_load(img)
    try:
        seams = ctx.carver_resize_width(n)
        print(w, h, n, "equal:", np.array_equal(seams, want["seams"]), "first diff seam", next((i for i in range(n) if not np.array_equal(seams[i], want["seams"][i])), -1))
        bad = next((i for i in range(n) if not np.array_equal(seams[i], want["seams"][i])), -1)
        if bad >= 0: print(seams[bad], want["seams"][bad])
    except dc.DctcError as e:
        print(w, h, n, "FAILED cuda", dc.lib().dctc_last_cuda_error(ctx.handle))
        break
