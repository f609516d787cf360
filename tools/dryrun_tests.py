"""Plan (not run) every graph of tests/test_ops_gpu.py on the engine in dry-run mode (no GPU).
Catches planner aborts / unsupported layouts before GPU time is spent."""
import os, sys, inspect
os.environ["GGML_B200_DRYRUN"] = "1"
os.environ.setdefault("GGML_B200_QUIET", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mlimgsynth_b200, blocks
import test_ops_gpu as T

eng = mlimgsynth_b200.load_engine(); eng.init_backend()
def fake_check(build, ref, e, tol, seed=0):
    G = blocks.Graph(eng); o = build(blocks.B(G, seed)); G.run(o); G.free(); return 0.0
T.check = fake_check
n = 0
for name, fn in inspect.getmembers(T, inspect.isfunction):
    if not name.startswith("test_"): continue
    marks = [m for m in getattr(fn, "pytestmark", []) if m.name == "parametrize"]
    cases = [()]
    if marks:
        cases = [c if isinstance(c, tuple) else (c,) for c in marks[0].args[1]]
    for c in cases:
        if name in ("test_multi_compute_and_reupload", "test_conv3x3_small_direct") or "monkeypatch" in inspect.signature(fn).parameters:
            continue
        print(name, c, flush=True)
        fn(eng, eng, *c); n += 1
print("planned", n, "graphs")
