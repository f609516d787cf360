"""Where one generation's time goes (SD1.5 512x512, 8 images, cfg 7, conditioning resident): device time (engine-stream CUDA events) and
wall time of  (a) the full generation  (b) the same without decode  (c) 1 and 2 sampler steps without decode (fixed cost + per-step cost).
Usage: phase_times.py [batch]"""
import os, sys, time, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench
from mlimgsynth_b200 import api
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
os.environ["GGML_B200_QUIET"] = "1"
ctx = api.Ctx(model=bench.weights_path("sd1"), image_dim=(512, 512), method="euler", cfg_scale=7, batch_size=B)
ctx.setup()
eng = bench.Engine()

def run(steps, no_decode, reps=3):
    ctx.set("steps", steps); ctx.set("no_decode", int(no_decode))
    out = []
    for i in range(reps + 2):
        ctx.set("seed", 1000 + i); ctx.set("prompt", bench.PROMPT)
        if i > 0:
            ctx.set("tensor_use_flags", api.TUF_CONDITIONING)
        eng.timer_start(); t0 = time.perf_counter()
        ctx.generate()
        dev = eng.timer_stop_ms(); wall = (time.perf_counter() - t0) * 1e3
        if i >= 2:
            out.append((dev, wall))
    return np.median([o[0] for o in out]), np.median([o[1] for o in out])

full = run(20, False); nodec = run(20, True); s1 = run(1, True); s2 = run(2, True); s1d = run(1, False)
print("full generation        : device %.1f ms, wall %.1f ms" % full)
print("20 steps, no decode    : device %.1f ms, wall %.1f ms  -> decode + pack + D2H in situ %.1f ms" % (nodec[0], nodec[1], full[0] - nodec[0]))
print("1 step, no decode      : device %.1f ms, wall %.1f ms" % s1)
print("2 steps, no decode     : device %.1f ms, wall %.1f ms  -> per step %.2f ms, fixed %.2f ms" % (s2[0], s2[1], s2[0] - s1[0], 2 * s1[0] - s2[0]))
print("1 step with decode     : device %.1f ms, wall %.1f ms  -> decode after a cool GPU %.1f ms" % (s1d[0], s1d[1], s1d[0] - s1[0]))
