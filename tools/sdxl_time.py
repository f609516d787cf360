import os, sys, time
ROOT = os.getcwd(); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench
from mlimgsynth_b200 import api
os.environ["GGML_B200_QUIET"] = "1"
cx = api.Ctx(model=bench.weights_path("sdxl"), image_dim=(1024, 1024), steps=4, method="euler", cfg_scale=7, batch_size=2)
marks = []
def cb(p):
    marks.append((time.perf_counter(), p.stage, p.step))
cx.set_callback(cb)
for i in range(4):
    marks.clear()
    cx.set("seed", 100 + i); cx.set("prompt", bench.PROMPT + (" %d" % i))
    t0 = time.perf_counter(); cx.generate(); t1 = time.perf_counter()
    imgs = [cx.image(k) for k in range(2)]; t2 = time.perf_counter()
    stages = {}
    prev = t0
    for t, st, step in marks:
        stages[st] = stages.get(st, 0) + (t - prev); prev = t
    print("gen %d: generate %.3f s, image get %.3f s, by stage (time until callback) %s, tail %.3f" % (i, t1 - t0, t2 - t1, {k: round(v, 3) for k, v in stages.items()}, t1 - prev))
