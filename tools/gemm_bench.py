"""Times single ops through the C ABI (CUDA events via the engine profile). Usage:
gemm_bench.py M,N,K [M,N,K ...]   (linear with bias);  conv:W,H,Cin,Cout,N for 3x3 convolutions;
geglu:M,K,D (GEGLU projection K -> 2D, gated to D);  res:M,N,K (linear + residual add);
gn:W,H,C,N (GroupNorm 32 + SiLU);  cgn:W,H,Cin,Cout,N (3x3 convolution -> GroupNorm 32 + SiLU);  ln:M,C (LayerNorm)."""
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
    elif spec.startswith("geglu:"):
        M, K, D = [int(x) for x in spec[6:].split(",")]
        y = b.geglu(b.inp(M, K, dtype=np.float16), D); flops = 2.0 * M * 2 * D * K
    elif spec.startswith("res:"):
        M, N, K = [int(x) for x in spec[4:].split(",")]
        x = b.inp(M, K, dtype=np.float16)
        y = G.g.ggml_add(G.cc, b.linear(x, N), b.inp(M, N, dtype=np.float16)); flops = 2.0 * M * N * K
    elif spec.startswith("gn:"):
        W, H, Ci, N = [int(x) for x in spec[3:].split(",")]
        x = b.conv2d(b.inp(N, 8, H, W), Ci, k=1, p=0)          # producer: activations arrive as f16 channels-last
        y = G.g.ggml_silu_inplace(G.cc, b.groupnorm32(x)); y = b.conv2d(y, 8, k=1, p=0); flops = 0.0
    elif spec.startswith("cgn:"):
        W, H, Ci, Co, N = [int(x) for x in spec[4:].split(",")]
        x = b.conv2d(b.inp(N, Ci, H, W), Co)                    # 3x3 producer of the group_norm: its epilogue takes the statistics
        y = G.g.ggml_silu_inplace(G.cc, b.groupnorm32(x)); y = b.conv2d(y, 8, k=1, p=0); flops = 2.0 * W * H * N * Co * Ci * 9
    elif spec.startswith("ln:"):
        M, Ci = [int(x) for x in spec[3:].split(",")]
        y = b.linear(b.layer_norm(b.linear(b.inp(M, 64, dtype=np.float16), Ci)), 64); flops = 0.0
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
    kinds = (8,) if spec.startswith("gn:") else (9,) if spec.startswith("ln:") else (14, 15)
    gb = 0.0
    for kind in kinds:
        a, b_, c_, l = C.c_double(), C.c_double(), C.c_double(), C.c_uint64()
        lib.ggml_b200_profile_get(kind, C.byref(a), C.byref(b_), C.byref(c_), C.byref(l))
        kms += a.value; gb += c_.value / 1e9
    kms /= 5; gb /= 5
    if spec.startswith("cgn:"):
        t = {}
        for kind in (15, 8):
            a, b_, c_, l = C.c_double(), C.c_double(), C.c_double(), C.c_uint64()
            lib.ggml_b200_profile_get(kind, C.byref(a), C.byref(b_), C.byref(c_), C.byref(l)); t[kind] = (a.value / 5 * 1e3, l.value // 5)
        lib.ggml_b200_profile_enable(0)
        print("%-28s conv %8.1f us (%d launch) | group_norm %8.1f us (%d launches) | graph %9.1f us" % (spec, t[15][0], t[15][1], t[8][0], t[8][1], ms * 1e3))
        G.free(); continue
    if not flops:
        print("%-28s kernel %8.1f us %7.2f TB/s (algorithmic bytes) | graph %9.1f us" % (spec, kms * 1e3, gb / kms if kms else 0, ms * 1e3))
        G.free(); continue
    lib.ggml_b200_profile_enable(0)
    print("%-28s kernel %8.1f us %7.1f TFLOP/s | graph incl. layout conversion %9.1f us" % (spec, kms * 1e3, flops / kms / 1e9 if kms else 0, ms * 1e3))
    G.free()
