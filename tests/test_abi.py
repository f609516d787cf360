"""The drop-in boundary: every symbol include/*.h declares is exported by the product library
(and by the oracle, which implements the same ABI). No compute calls here (runs without a GPU)."""
import ctypes, os, re
import pytest
import mlimgsynth_b200
from mlimgsynth_b200.ggml import ABI_SYMBOLS

ROOT = mlimgsynth_b200.ROOT


def declared_symbols():
    names = set()
    for h in ("ggml.h", "ggml-alloc.h", "ggml-backend.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        for m in re.finditer(r"GGML_API[^;]*?\b(ggml_[a-z0-9_]+)\s*\(", src, flags=re.S):
            names.add(m.group(1))
    return names


def test_header_symbol_inventory():
    decl = declared_symbols()
    assert decl == set(ABI_SYMBOLS), (decl ^ set(ABI_SYMBOLS))
    assert len(decl) >= 86


def test_engine_exports_every_declared_symbol():
    lib = ctypes.CDLL(mlimgsynth_b200.engine_path())
    missing = [s for s in sorted(declared_symbols()) if not hasattr(lib, s)]
    assert not missing, missing


def test_oracle_exports_every_declared_symbol(oracle_built):
    lib = ctypes.CDLL(os.path.join(oracle_built, "libggml_ref.so"))
    missing = [s for s in sorted(declared_symbols()) if not hasattr(lib, s)]
    assert not missing, missing


def test_engine_metadata_calls_without_gpu():
    """Pure host-side entry points work with no device: type tables, fp16 conversion, recording."""
    from mlimgsynth_b200.ggml import GGML, Graph
    import numpy as np
    g = GGML(mlimgsynth_b200.engine_path())
    g.lib.ggml_type_size.restype = ctypes.c_size_t
    assert g.lib.ggml_type_size(0) == 4 and g.lib.ggml_type_size(1) == 2
    g.lib.ggml_fp16_to_fp32.restype = ctypes.c_float
    g.lib.ggml_fp16_to_fp32.argtypes = [ctypes.c_uint16]
    assert g.lib.ggml_fp16_to_fp32(0x3C00) == 1.0
    G = Graph(g)
    x = G.leaf(np.zeros((1, 8, 6, 5), np.float32))
    w = G.leaf(np.zeros((16, 8, 3, 3), np.float16))
    y = g.ggml_conv_2d(G.cc, w, x, 2, 2, 1, 1, 1, 1)
    assert G.shape(y) == (3, 3, 16, 1)
    p = g.ggml_permute(G.cc, y, 1, 2, 0, 3)
    assert G.shape(p) == (16, 3, 3, 1)
    G.free()


def test_no_cpu_backend():
    from mlimgsynth_b200.ggml import GGML
    g = GGML(mlimgsynth_b200.engine_path())
    assert not g.lib.ggml_backend_init_by_name(b"CPU", None)
