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


# Every function the reference's public header declares (/root/reference/include/mlimgsynth.h:395-558; mlis_ctx_create and
# mlis_tensor_for are macros, mlis_text_cond_encode is commented out there, mlis_state_str appears only in a comment).
REFERENCE_MLIS_API = """mlis_backend_info_get mlis_clip_text_encode mlis_ctx_create_i mlis_ctx_destroy mlis_errstr_get mlis_generate
mlis_image_decode mlis_image_encode mlis_image_get mlis_infotext_get mlis_loglvl_fromz mlis_loglvl_str mlis_mask_encode
mlis_method_fromz mlis_method_str mlis_model_type_desc mlis_model_type_fromz mlis_model_type_str mlis_option_fromz
mlis_option_get mlis_option_set mlis_option_set_str mlis_option_str mlis_sched_fromz mlis_sched_str mlis_setup mlis_stage_desc
mlis_stage_fromz mlis_stage_str mlis_tensor_copy mlis_tensor_count mlis_tensor_free mlis_tensor_get mlis_tensor_resize
mlis_tensor_resize_like mlis_tensor_similarity mlis_text_tokenize""".split()


def host_header_symbols():
    src = open(os.path.join(ROOT, "include", "mlimgsynth_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"\b(mlis_[a-z0-9_]+)\s*\(", src)) - {"mlis_ctx_create"}


def test_host_library_exports_the_reference_public_api():
    """Level-2 boundary: libmlimgsynth_b200.so exports every function of the reference's include/mlimgsynth.h and
    everything include/mlimgsynth_b200.h declares on top of it."""
    lib = ctypes.CDLL(mlimgsynth_b200.HOST_LIB)
    decl = host_header_symbols()
    assert set(REFERENCE_MLIS_API) <= decl, set(REFERENCE_MLIS_API) - decl
    missing = [s for s in sorted(decl) if not hasattr(lib, s)]
    assert not missing, missing


def test_reference_header_inventory_is_current():
    """Where the reference tree is present (build container), re-derive the list above from its header."""
    h = "/root/reference/include/mlimgsynth.h"
    if not os.path.exists(h):
        pytest.skip("reference tree not on this machine")
    src = re.sub(r"//[^\n]*", "", re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S))
    names = set(re.findall(r"\b(mlis_[a-z0-9_]+)\s*\(", src)) - {"mlis_ctx_create", "mlis_tensor_for"}
    assert names == set(REFERENCE_MLIS_API), names ^ set(REFERENCE_MLIS_API)


def test_enum_string_functions_match_reference_tables():
    """mlimgsynth.c:212-302 name/description tables (the five functions missing in round 1 included)."""
    L = ctypes.CDLL(mlimgsynth_b200.HOST_LIB)
    for f in ("mlis_stage_str", "mlis_stage_desc", "mlis_loglvl_str", "mlis_model_type_desc", "mlis_model_type_str"):
        getattr(L, f).restype = ctypes.c_char_p
    assert [L.mlis_stage_desc(i) for i in range(5)] == [b"Idle", b"Conditioning encoding", b"Image encoding", b"Image decoding", b"Denoising"]
    assert L.mlis_stage_desc(9) == b"???"
    assert [L.mlis_stage_fromz(s) for s in (b"idle", b"cond_encode", b"image-decode", b"denoise", b"nope")] == [0, 1, 3, 4, -1]
    assert [L.mlis_loglvl_fromz(s) for s in (b"none", b"error", b"warning", b"info", b"verbose", b"debug", b"max", b"x")] == [0, 10, 20, 30, 40, 50, 255, -1]
    assert L.mlis_loglvl_str(40) == b"verbose" and L.mlis_loglvl_str(41) == b"???"
    assert [L.mlis_model_type_desc(i) for i in range(4)] == [b"None", b"Stable Diffusion 1.x", b"Stable Diffusion 2.x", b"Stable Diffusion XL"]
    assert L.mlis_option_fromz(b"CFG_SCALE") == 12 and L.mlis_option_fromz(b"cfg-scale") == 12 and L.mlis_option_fromz(b"bogus") == -1
    assert L.mlis_method_fromz(b"dpm++2m") == 4
