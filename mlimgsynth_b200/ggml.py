"""ctypes binding of the ggml-shaped C ABI (include/ggml.h, ggml-alloc.h, ggml-backend.h).

The same binding drives the product engine (lib/libggml_b200.so) and -- in tests only -- the CPU
oracle (oracle/_ref/libggml_ref.so), so a parity test builds one graph description and runs it on
both. Nothing here computes; it mirrors how the reference's C code calls the boundary.
"""
import ctypes as C
import numpy as np

GGML_TYPE_F32, GGML_TYPE_F16, GGML_TYPE_I32 = 0, 1, 26
SCALE_NEAREST = 0

c_tensor_p = C.c_void_p


class Tensor(C.Structure):
    _fields_ = [
        ("type", C.c_int),
        ("buffer", C.c_void_p),
        ("ne", C.c_int64 * 4),
        ("nb", C.c_size_t * 4),
        ("op", C.c_int),
        ("op_params", C.c_int32 * 16),
        ("flags", C.c_int32),
        ("src", C.c_void_p * 10),
        ("view_src", C.c_void_p),
        ("view_offs", C.c_size_t),
        ("data", C.c_void_p),
        ("name", C.c_char * 64),
        ("extra", C.c_void_p),
        ("padding", C.c_char * 8),
    ]


class InitParams(C.Structure):
    _fields_ = [("mem_size", C.c_size_t), ("mem_buffer", C.c_void_p), ("no_alloc", C.c_bool)]


_P = C.c_void_p
_SIGS = {
    # name: (restype, argtypes)
    "ggml_init": (_P, [InitParams]),
    "ggml_free": (None, [_P]),
    "ggml_new_tensor_1d": (_P, [_P, C.c_int, C.c_int64]),
    "ggml_new_tensor_2d": (_P, [_P, C.c_int, C.c_int64, C.c_int64]),
    "ggml_new_tensor_4d": (_P, [_P, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int64]),
    "ggml_new_graph_custom": (_P, [_P, C.c_size_t, C.c_bool]),
    "ggml_build_forward_expand": (None, [_P, _P]),
    "ggml_graph_n_nodes": (C.c_int, [_P]),
    "ggml_set_name": (_P, [_P, C.c_char_p]),
    "ggml_set_input": (None, [_P]),
    "ggml_set_output": (None, [_P]),
    "ggml_nbytes": (C.c_size_t, [_P]),
    "ggml_nelements": (C.c_int64, [_P]),
    "ggml_add": (_P, [_P, _P, _P]),
    "ggml_add_inplace": (_P, [_P, _P, _P]),
    "ggml_mul": (_P, [_P, _P, _P]),
    "ggml_scale": (_P, [_P, _P, C.c_float]),
    "ggml_scale_inplace": (_P, [_P, _P, C.c_float]),
    "ggml_mul_mat": (_P, [_P, _P, _P]),
    "ggml_conv_2d": (_P, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "ggml_norm": (_P, [_P, _P, C.c_float]),
    "ggml_group_norm": (_P, [_P, _P, C.c_int, C.c_float]),
    "ggml_silu": (_P, [_P, _P]),
    "ggml_silu_inplace": (_P, [_P, _P]),
    "ggml_gelu_inplace": (_P, [_P, _P]),
    "ggml_gelu_quick_inplace": (_P, [_P, _P]),
    "ggml_relu_inplace": (_P, [_P, _P]),
    "ggml_tanh_inplace": (_P, [_P, _P]),
    "ggml_soft_max_inplace": (_P, [_P, _P]),
    "ggml_diag_mask_inf_inplace": (_P, [_P, _P, C.c_int]),
    "ggml_cont": (_P, [_P, _P]),
    "ggml_permute": (_P, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int]),
    "ggml_transpose": (_P, [_P, _P]),
    "ggml_reshape_3d": (_P, [_P, _P, C.c_int64, C.c_int64, C.c_int64]),
    "ggml_reshape_4d": (_P, [_P, _P, C.c_int64, C.c_int64, C.c_int64, C.c_int64]),
    "ggml_view_1d": (_P, [_P, _P, C.c_int64, C.c_size_t]),
    "ggml_view_4d": (_P, [_P, _P, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t]),
    "ggml_concat": (_P, [_P, _P, _P, C.c_int]),
    "ggml_pad": (_P, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int]),
    "ggml_upscale": (_P, [_P, _P, C.c_int, C.c_int]),
    "ggml_get_rows": (_P, [_P, _P, _P]),
    "ggml_timestep_embedding": (_P, [_P, _P, C.c_int, C.c_int]),
    "ggml_gallocr_new": (_P, [_P]),
    "ggml_gallocr_free": (None, [_P]),
    "ggml_gallocr_reserve": (C.c_bool, [_P, _P]),
    "ggml_gallocr_alloc_graph": (C.c_bool, [_P, _P]),
    "ggml_gallocr_get_buffer_size": (C.c_size_t, [_P, C.c_int]),
    "ggml_backend_init_by_name": (_P, [C.c_char_p, C.c_char_p]),
    "ggml_backend_init_best": (_P, []),
    "ggml_backend_free": (None, [_P]),
    "ggml_backend_name": (C.c_char_p, [_P]),
    "ggml_backend_get_default_buffer_type": (_P, [_P]),
    "ggml_backend_graph_compute": (C.c_int, [_P, _P]),
    "ggml_backend_tensor_set": (None, [_P, _P, C.c_size_t, C.c_size_t]),
    "ggml_backend_tensor_get": (None, [_P, _P, C.c_size_t, C.c_size_t]),
    "ggml_backend_reg_count": (C.c_size_t, []),
    "ggml_backend_reg_get": (_P, [C.c_size_t]),
    "ggml_backend_reg_name": (C.c_char_p, [_P]),
    "ggml_backend_reg_get_proc_address": (_P, [_P, C.c_char_p]),
}

# every symbol include/*.h declares (checked by tests/test_abi.py against both libraries)
ABI_SYMBOLS = """
ggml_abort ggml_init ggml_free ggml_tensor_overhead ggml_graph_overhead ggml_new_tensor_1d ggml_new_tensor_2d
ggml_new_tensor_4d ggml_new_graph_custom ggml_build_forward_expand ggml_graph_size ggml_graph_n_nodes
ggml_get_first_tensor ggml_get_next_tensor ggml_set_name ggml_get_name ggml_set_input ggml_set_output ggml_nbytes
ggml_nelements ggml_element_size ggml_type_size ggml_type_name ggml_n_dims ggml_op_name ggml_op_desc
ggml_get_type_traits ggml_add ggml_add_inplace ggml_mul ggml_scale ggml_scale_inplace ggml_mul_mat ggml_conv_2d
ggml_norm ggml_group_norm ggml_silu ggml_silu_inplace ggml_gelu_inplace ggml_gelu_quick_inplace ggml_relu_inplace
ggml_tanh_inplace ggml_soft_max_inplace ggml_diag_mask_inf_inplace ggml_cont ggml_permute ggml_transpose
ggml_reshape_3d ggml_reshape_4d ggml_view_1d ggml_view_4d ggml_concat ggml_pad ggml_upscale ggml_get_rows
ggml_timestep_embedding ggml_map_custom1_inplace ggml_fp16_to_fp32 ggml_fp32_to_fp16 ggml_fp16_to_fp32_row
ggml_fp32_to_fp16_row ggml_bf16_to_fp32_row ggml_quantize_chunk
ggml_gallocr_new ggml_gallocr_free ggml_gallocr_reserve ggml_gallocr_alloc_graph ggml_gallocr_get_buffer_size
ggml_backend_init_by_name ggml_backend_init_best ggml_backend_free ggml_backend_name
ggml_backend_get_default_buffer_type ggml_backend_get_device ggml_backend_graph_compute ggml_backend_tensor_set
ggml_backend_tensor_get ggml_backend_buffer_is_host ggml_backend_reg_count ggml_backend_reg_get
ggml_backend_reg_name ggml_backend_reg_dev_count ggml_backend_reg_dev_get ggml_backend_reg_get_proc_address
ggml_backend_dev_name ggml_backend_dev_description ggml_backend_dev_memory ggml_backend_dev_backend_reg
""".split()


class Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("graph_launches", C.c_uint64), ("plans_built", C.c_uint64),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64)]


class GGML:
    """One loaded implementation of the boundary."""

    def __init__(self, path):
        self.path = path
        self.lib = C.CDLL(path, mode=C.RTLD_LOCAL)
        for name, (res, args) in _SIGS.items():
            f = getattr(self.lib, name)
            f.restype, f.argtypes = res, args
        self.backend = None

    def __getattr__(self, name):
        if name.startswith("ggml_"):
            return getattr(self.lib, name)
        raise AttributeError(name)

    def init_backend(self, name=None):
        if self.backend is None:
            self.backend = self.lib.ggml_backend_init_by_name(name.encode() if name else None, None)
            if not self.backend:
                raise RuntimeError("backend init failed for %s" % self.path)
        return self.backend

    def stats(self):
        reg = self.lib.ggml_backend_reg_get(0)
        fn = self.lib.ggml_backend_reg_get_proc_address(reg, b"ggml_b200_stats")
        if not fn:
            return None
        st = C.cast(C.CFUNCTYPE(C.POINTER(Stats))(fn)(), C.POINTER(Stats)).contents
        return {k: getattr(st, k) for k, _ in Stats._fields_}


_NP = {GGML_TYPE_F32: np.float32, GGML_TYPE_F16: np.float16, GGML_TYPE_I32: np.int32}


class Graph:
    """Builds and runs one graph the way mlblock.c does: params context + compute context,
    build_forward_expand, gallocr, tensor_set, graph_compute, tensor_get."""

    def __init__(self, g: GGML, n_max=16384):
        self.g = g
        ip = InitParams(1 << 20, None, True)
        self.cp = g.ggml_init(ip)
        self.cc = g.ggml_init(ip)
        self.n_max = n_max
        self.leaves = []          # (tensor, numpy array)
        self.graph = None
        self.alloc = None
        self.outputs = []

    # ---- leaves
    def leaf(self, arr, name=None):
        """New leaf holding `arr` (numpy shape is REVERSED ggml shape, i.e. C-order)."""
        arr = np.ascontiguousarray(arr)
        tp = {np.dtype(np.float32): GGML_TYPE_F32, np.dtype(np.float16): GGML_TYPE_F16, np.dtype(np.int32): GGML_TYPE_I32}[arr.dtype]
        ne = list(arr.shape[::-1]) + [1] * (4 - arr.ndim)
        t = self.g.ggml_new_tensor_4d(self.cp, tp, *ne)
        if name:
            self.g.ggml_set_name(t, name.encode())
        self.leaves.append((t, arr))
        return t

    def t(self, ptr):
        return C.cast(ptr, C.POINTER(Tensor)).contents

    def shape(self, ptr):
        return tuple(self.t(ptr).ne)

    def nb(self, ptr):
        return tuple(self.t(ptr).nb)

    # ---- run
    def build(self, *outputs):
        g = self.g
        self.outputs = list(outputs)
        for o in outputs:
            g.ggml_set_output(o)
        self.graph = g.ggml_new_graph_custom(self.cc, self.n_max, False)
        for o in outputs:
            g.ggml_build_forward_expand(self.graph, o)
        be = g.init_backend()
        self.alloc = g.ggml_gallocr_new(g.ggml_backend_get_default_buffer_type(be))
        assert g.ggml_gallocr_reserve(self.alloc, self.graph)
        assert g.ggml_gallocr_alloc_graph(self.alloc, self.graph)
        for t, arr in self.leaves:
            if self.t(t).data:
                g.ggml_backend_tensor_set(t, arr.ctypes.data_as(C.c_void_p), 0, arr.nbytes)

    def set(self, t, arr):
        arr = np.ascontiguousarray(arr)
        self.g.ggml_backend_tensor_set(t, arr.ctypes.data_as(C.c_void_p), 0, arr.nbytes)

    def compute(self):
        r = self.g.ggml_backend_graph_compute(self.g.backend, self.graph)
        if r != 0:
            raise RuntimeError("graph compute failed: %d" % r)

    def get(self, ptr):
        tt = self.t(ptr)
        ne = [int(x) for x in tt.ne]
        out = np.empty(ne[::-1], dtype=_NP[tt.type])
        assert out.nbytes == self.g.ggml_nbytes(ptr), "output must be contiguous"
        self.g.ggml_backend_tensor_get(ptr, out.ctypes.data_as(C.c_void_p), 0, out.nbytes)
        return out

    def run(self, *outputs):
        self.build(*outputs)
        self.compute()
        return [self.get(o) for o in outputs]

    def free(self):
        if self.alloc:
            self.g.ggml_gallocr_free(self.alloc)
            self.alloc = None
        if self.cc:
            self.g.ggml_free(self.cc)
            self.g.ggml_free(self.cp)
            self.cc = self.cp = None
