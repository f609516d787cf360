"""mlimgsynth_b200 -- B200-native (sm_100a) execution engine for libmlimgsynth's denoising path.

The product is a C-ABI shared library (lib/libggml_b200.so) that exports the ggml-shaped
graph/alloc/backend interface declared in include/ggml*.h; this Python package only locates,
builds and binds it (ctypes) for the tests and the benchmark. There is no Python compute path
and no CPU fallback: if the CUDA library is missing, loading fails loudly.
"""
import os

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
ENGINE_LIB = os.path.join(HERE, "lib", "libggml_b200.so")
HOST_LIB = os.path.join(HERE, "lib", "libmlimgsynth_b200.so")


class EngineMissing(RuntimeError):
    pass


def engine_path():
    if not os.path.exists(ENGINE_LIB):
        raise EngineMissing(
            "CUDA engine %s not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback." % ENGINE_LIB)
    return ENGINE_LIB


def load_engine():
    from .ggml import GGML
    return GGML(engine_path())
