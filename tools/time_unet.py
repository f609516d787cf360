"""Graph-replay time of one batched UNet evaluation through the public API (wall clock over 20 calls after warm-up; the
latent / conditioning uploads are inside, the same for every variant). Usage: time_unet.py [batch] [sd1|sd2|sdxl]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench
from mlimgsynth_b200 import api
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 16
kind = sys.argv[2] if len(sys.argv) > 2 else "sd1"
side = {"sd1": 64, "sd2": 96, "sdxl": 128}[kind]; nctx = {"sd1": 768, "sd2": 1024, "sdxl": 2048}[kind]
os.environ["GGML_B200_QUIET"] = "1"
ctx = api.Ctx(model=bench.weights_path(kind))
x = np.random.default_rng(0).standard_normal((nb, 4, side, side)).astype(np.float32)
label = (np.random.default_rng(2).standard_normal((nb, 2816)) * 0.5).astype(np.float32) if kind == "sdxl" else None
cond = (np.random.default_rng(1).standard_normal((nb, 77, nctx)) * 0.5).astype(np.float32)
for _ in range(5): y = ctx.unet_eval(x, cond, label, 5.0)
ts = []
for _ in range(4):
    t0 = time.perf_counter()
    for _ in range(10): y = ctx.unet_eval(x, cond, label, 5.0)
    ts.append((time.perf_counter() - t0) / 10 * 1e3)
print("%s batch %d: unet_eval %.3f ms (min of 4 x 10 calls; all %s) checksum %.6e" % (kind, nb, min(ts), " ".join("%.3f" % t for t in ts), float(np.abs(y).sum())))
