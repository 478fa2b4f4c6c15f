import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
import dct_carver_b200 as dc, oracle_lib as ol
if os.environ.get("DCTC_LIB"): dc.LIB_PATH = os.environ["DCTC_LIB"]
ctx = dc.Context(0)
ctx.set_params(8, 0.5, 0.5)
for (w, h) in [(128, 16), (128, 24), (128, 32), (128, 40), (256, 64), (200, 150)]:
    img = ol.synth_image(w, h, 3, 5, 0)
    ctx.set_kernel(dc.KERNEL_TC_SPLIT)
    try:
        got = ctx.energy_full(img)
    except dc.DctcError as e:
        print(w, h, "FAILED cuda", dc.lib().dctc_last_cuda_error(ctx.handle)); break
    want = ol.oracle_energy(img, 8, 0.5, 0.5)
    print(w, h, "max err", np.abs(got - want).max())
