"""First-light diagnostics on the GPU box: each case in its own subprocess (a trapped kernel kills
only that case), TC path and SIMT path side by side against the oracle."""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASES = {
    "linear_128x64x64": "b.linear(b.inp(128, 64), 64, bias=False)",
    "linear_128x64x128": "b.linear(b.inp(128, 128), 64, bias=False)",
    "linear_256x128x320": "b.linear(b.inp(256, 320), 128)",
    "linear_77x320x768": "b.linear(b.inp(77, 768), 320)",
    "linear_4096x320x320": "b.linear(b.inp(4096, 320), 320)",
    "conv1x1_32x32_320": "b.conv2d(b.inp(1, 320, 32, 32), 640, 1, 1, 0)",
    "conv3x3_16x16_64": "b.conv2d(b.inp(1, 64, 16, 16), 64)",
    "conv3x3_64x64_320": "b.conv2d(b.inp(1, 320, 64, 64), 320)",
    "conv3x3_8x8_1280_n2": "b.conv2d(b.inp(2, 1280, 8, 8), 1280)",
    "conv_in_4": "b.conv2d(b.inp(1, 4, 64, 64), 320)",
    "groupnorm": "b.g.ggml_silu_inplace(b.cc, b.groupnorm32(b.inp(1, 320, 32, 32, scale=2.0)))",
    "layernorm": "b.layer_norm(b.inp(100, 320, scale=3.0))",
    "attn": "b.attn_mhead(b.inp(256, 320), None, None, 320, 320, 8)",
    "resnet": "b.resnet(b.inp(1, 320, 32, 32), b.inp(1, 1280), 640)",
}

CHILD = r'''
import os, sys, numpy as np
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import mlimgsynth_b200
from mlimgsynth_b200.ggml import GGML
from blocks import run_both, max_rel_err, B
eng = mlimgsynth_b200.load_engine(); eng.init_backend()
ref = GGML(os.path.join(%(root)r, "oracle", "_ref", "libggml_ref.so"))
def build(b):
    _orig = b.attn_mhead
    def am(q, k, v, *a, **kw):
        return _orig(q, q if k is None else k, q if v is None else v, *a, **kw)
    b.attn_mhead = am
    return %(expr)s
(r,), (e,) = run_both(build, ref, eng, 0)
d = np.abs(e - r)
print("RESULT err=%%.3e  max|ref|=%%.3g  nan=%%d  badfrac=%%.4f  first_bad=%%s" %% (max_rel_err(e, r), np.abs(r).max(), int(np.isnan(e).sum()),
      float((d > 1e-2 * np.abs(r).max()).mean()), str(np.argwhere(d > 1e-2 * np.abs(r).max())[:3].tolist())))
'''

def main():
    sel = sys.argv[1:] or list(CASES)
    for name in sel:
        for mode in ("tc", "simt"):
            env = dict(os.environ, GGML_B200_SYNC="1", GGML_B200_QUIET="1")
            if mode == "simt":
                env["GGML_B200_GEMM"] = "simt"
            code = CHILD % {"root": ROOT, "expr": CASES[name]}
            try:
                p = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=180)
                out = [l for l in p.stdout.splitlines() if l.startswith("RESULT")]
                msg = out[0] if out else "FAILED rc=%d: %s" % (p.returncode, (p.stderr or p.stdout)[-400:].replace("\n", " | "))
            except subprocess.TimeoutExpired:
                msg = "TIMEOUT"
            print("%-24s %-5s %s" % (name, mode, msg), flush=True)

if __name__ == "__main__":
    main()
