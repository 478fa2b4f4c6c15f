import os, sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import dct_carver_b200 as dc, oracle_lib as ol
if os.environ.get("DCTC_LIB"): dc.LIB_PATH = os.environ["DCTC_LIB"]
img = ol.synth_image(3840, 2160, 3, 0xD0C7CA14, 0)
ctx = dc.Context(0); ctx.set_params(8, 0.5, 0.5); ctx.carver_load(img); ctx.carver_resize_width(4)
t = time.perf_counter(); ctx.carver_resize_width(60); dt = time.perf_counter() - t
print("3840x2160: %.1f us per seam" % (dt / 60 * 1e6))
