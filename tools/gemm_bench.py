"""Times single linear / conv shapes through the C ABI (CUDA events via the engine profile). Usage:
gemm_bench.py M,N,K [M,N,K ...]   (linear with bias);  conv:W,H,Cin,Cout,N for 3x3 convolutions."""
import os, sys, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("GGML_B200_QUIET", "1")
import mlimgsynth_b200
from mlimgsynth_b200.ggml import Graph
from blocks import B
eng = mlimgsynth_b200.load_engine(); eng.init_backend()
lib = C.CDLL(mlimgsynth_b200.ENGINE_LIB)
lib.ggml_b200_timer_stop.restype = C.c_double
for spec in sys.argv[1:]:
    G = Graph(eng); b = B(G, 0)
    if spec.startswith("conv:"):
        W, H, Ci, Co, N = [int(x) for x in spec[5:].split(",")]
        y = b.conv2d(b.inp(N, Ci, H, W), Co); flops = 2.0 * W * H * N * Co * Ci * 9
    else:
        M, N, K = [int(x) for x in spec.split(",")]
        y = b.linear(b.inp(M, K, dtype=np.float16), N); flops = 2.0 * M * N * K
    G.build(y)
    for _ in range(3): G.compute()
    lib.ggml_b200_timer_start()
    n = 20
    for _ in range(n): G.compute()
    ms = lib.ggml_b200_timer_stop() / n
    # kernel-only time: profiled eager pass (CUDA events around each step), tcgen05 GEMM + conv kinds
    lib.ggml_b200_profile_enable(1)
    for _ in range(5): G.compute()
    kms = 0.0
    for kind in (14, 15):
        a, b_, c_, l = C.c_double(), C.c_double(), C.c_double(), C.c_uint64()
        lib.ggml_b200_profile_get(kind, C.byref(a), C.byref(b_), C.byref(c_), C.byref(l))
        kms += a.value
    kms /= 5
    lib.ggml_b200_profile_enable(0)
    print("%-28s kernel %8.1f us %7.1f TFLOP/s | graph incl. layout conversion %9.1f us" % (spec, kms * 1e3, flops / kms / 1e9 if kms else 0, ms * 1e3))
    G.free()
