"""Per-step CUDA-event times of one batched UNet evaluation (eager, profiled). Usage: profile_unet.py [batch]"""
import os, sys, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench
from mlimgsynth_b200 import api
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 16
os.environ["GGML_B200_QUIET"] = "1"
ctx = api.Ctx(model=bench.weights_path("sd1"))
eng = C.CDLL(os.path.join(ROOT, "mlimgsynth_b200", "lib", "libggml_b200.so"))
x = np.random.default_rng(0).standard_normal((nb, 4, 64, 64)).astype(np.float32)
cond = (np.random.default_rng(1).standard_normal((nb, 77, 768)) * 0.5).astype(np.float32)
for _ in range(3): ctx.unet_eval(x, cond, None, 5.0)
os.environ["GGML_B200_PROFILE_STEPS"] = "1"
eng.ggml_b200_profile_enable(1)
ctx.unet_eval(x, cond, None, 5.0)
eng.ggml_b200_profile_enable(0)
