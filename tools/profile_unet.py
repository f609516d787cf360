"""Per-step CUDA-event times of one batched UNet evaluation (eager, profiled). Usage: profile_unet.py [batch]"""
import os, sys, ctypes as C
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
eng = C.CDLL(os.path.join(ROOT, "mlimgsynth_b200", "lib", "libggml_b200.so"))
x = np.random.default_rng(0).standard_normal((nb, 4, side, side)).astype(np.float32)
label = (np.random.default_rng(2).standard_normal((nb, 2816)) * 0.5).astype(np.float32) if kind == "sdxl" else None
cond = (np.random.default_rng(1).standard_normal((nb, 77, nctx)) * 0.5).astype(np.float32)
for _ in range(3): ctx.unet_eval(x, cond, label, 5.0)
os.environ["GGML_B200_PROFILE_STEPS"] = "1"
eng.ggml_b200_profile_enable(1)
ctx.unet_eval(x, cond, label, 5.0)
eng.ggml_b200_profile_enable(0)
