"""Generate with a random-init checkpoint of a given family and report timings (development aid).
Usage: run_model.py KIND W H BATCH STEPS [METHOD]   e.g. run_model.py sdxl 1024 1024 2 4"""
import os, sys, time, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench
from mlimgsynth_b200 import api
kind, W, H, B, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
method = sys.argv[6] if len(sys.argv) > 6 else "euler"
os.environ.setdefault("GGML_B200_QUIET", "1")
t0 = time.time(); model = bench.weights_path(kind); print("weights %s: %.1f s" % (model, time.time() - t0), flush=True)
eng = C.CDLL(os.path.join(ROOT, "mlimgsynth_b200", "lib", "libggml_b200.so"), mode=C.RTLD_LOCAL)
eng.ggml_b200_timer_stop.restype = C.c_double
t0 = time.time()
ctx = api.Ctx(backend="B200:0", model=model, image_dim=(W, H), steps=steps, method=method, cfg_scale=7, batch_size=B)
ctx.set("prompt", bench.PROMPT); ctx.setup()
print("setup %.1f s" % (time.time() - t0), flush=True)
for i in range(3):
    ctx.set("seed", 42 + i); ctx.set("prompt", bench.PROMPT)
    t1 = time.time(); eng.ggml_b200_timer_start(); ctx.generate(); ms = eng.ggml_b200_timer_stop()
    img = ctx.image(0)
    print("generate %d: device %.1f ms, wall %.1f ms, image %s mean %.1f std %.1f" % (i, ms, (time.time() - t1) * 1e3, img.shape, img.mean(), img.std()), flush=True)
lat = ctx.tensor(api.TENSOR_LATENT)
print("latent", lat.shape, "finite", bool(np.isfinite(lat).all()), "std %.3f" % lat.std())
print("per sampler step (incl. decode amortised): %.1f ms ; images/s %.2f" % (ms / steps, B / (ms / 1e3)))
