"""Per-step CUDA-event times of one VAE decode (eager, profiled). Usage: profile_vae.py [n_images] [latent_side]"""
import os, sys, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench
from mlimgsynth_b200 import api
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 8
side = int(sys.argv[2]) if len(sys.argv) > 2 else 64
os.environ["GGML_B200_QUIET"] = "1"
ctx = api.Ctx(model=bench.weights_path("sd1"))
eng = C.CDLL(os.path.join(ROOT, "mlimgsynth_b200", "lib", "libggml_b200.so"))
eng.ggml_b200_timer_stop.restype = C.c_double
lat = (np.random.default_rng(0).standard_normal((nb, 4, side, side)) * 0.18215).astype(np.float32)
for _ in range(2): ctx.decode(lat)
eng.ggml_b200_timer_start()
ctx.decode(lat)
print("vae decode of %d latents %dx%d: %.2f ms (graph replay)" % (nb, side, side, eng.ggml_b200_timer_stop()), file=sys.stderr)
os.environ["GGML_B200_PROFILE_STEPS"] = "1"
eng.ggml_b200_profile_enable(1)
ctx.decode(lat)
eng.ggml_b200_profile_enable(0)
