"""ctypes binding of libmlimgsynth_b200.so (include/mlimgsynth_b200.h) -- the mlis_* API a user of
the reference's python/mlimgsynth.py would call. Used by bench.py and the tests."""
import ctypes as C
import os
import numpy as np
from . import HOST_LIB, ENGINE_LIB, EngineMissing


class MLIS_Image(C.Structure):
    _fields_ = [("d", C.POINTER(C.c_uint8)), ("sz", C.c_size_t), ("w", C.c_uint), ("h", C.c_uint), ("c", C.c_uint), ("flags", C.c_int)]


class MLIS_Tensor(C.Structure):
    _fields_ = [("d", C.POINTER(C.c_float)), ("n", C.c_int * 4), ("flags", C.c_int)]


class MLIS_Progress(C.Structure):
    _fields_ = [("stage", C.c_int), ("step", C.c_int), ("step_end", C.c_int), ("nfe", C.c_int), ("step_time", C.c_double), ("time", C.c_double)]


TENSOR_IMAGE, TENSOR_MASK, TENSOR_LATENT, TENSOR_LMASK, TENSOR_COND, TENSOR_LABEL, TENSOR_NCOND, TENSOR_NLABEL = range(1, 9)
TUF_IMAGE, TUF_MASK, TUF_LATENT, TUF_LMASK, TUF_CONDITIONING = 1, 2, 4, 8, 16
SUBMODEL_CLIP, SUBMODEL_CLIP2 = 4, 5
OPT_IMAGE, OPT_IMAGE_MASK, OPT_CALLBACK = 20, 21, 30
CALLBACK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.POINTER(MLIS_Progress))
CFG_EXCHANGE = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t)

_lib = None


def bind(L, extensions=True):
    """Declare the mlis_* prototypes (include/mlimgsynth_b200.h) on a loaded library. `extensions=False` binds only the
    functions of the reference's own header (any library with that ABI can then be driven by `Ctx(_lib=...)`)."""
    L.mlis_ctx_create_i.restype = C.c_void_p
    L.mlis_ctx_create_i.argtypes = [C.c_int]
    L.mlis_ctx_destroy.argtypes = [C.POINTER(C.c_void_p)]
    L.mlis_errstr_get.restype = C.c_char_p
    L.mlis_errstr_get.argtypes = [C.c_void_p]
    L.mlis_option_set_str.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
    L.mlis_generate.argtypes = [C.c_void_p]
    L.mlis_setup.argtypes = [C.c_void_p]
    L.mlis_image_get.restype = C.POINTER(MLIS_Image)
    L.mlis_image_get.argtypes = [C.c_void_p, C.c_int]
    L.mlis_infotext_get.restype = C.c_char_p
    L.mlis_infotext_get.argtypes = [C.c_void_p, C.c_int]
    L.mlis_tensor_get.restype = C.POINTER(MLIS_Tensor)
    L.mlis_tensor_get.argtypes = [C.c_void_p, C.c_int]
    L.mlis_text_tokenize.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.POINTER(C.c_int32)), C.c_int]
    L.mlis_clip_text_encode.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(MLIS_Tensor), C.POINTER(MLIS_Tensor), C.c_int, C.c_int]
    L.mlis_image_decode.argtypes = [C.c_void_p, C.POINTER(MLIS_Tensor), C.POINTER(MLIS_Tensor), C.c_int]
    L.mlis_image_encode.argtypes = [C.c_void_p, C.POINTER(MLIS_Tensor), C.POINTER(MLIS_Tensor), C.c_int]
    L.mlis_tensor_resize.argtypes = [C.POINTER(MLIS_Tensor), C.c_int, C.c_int, C.c_int, C.c_int]
    L.mlis_tensor_free.argtypes = [C.POINTER(MLIS_Tensor)]
    if extensions:
        L.mlis_unet_eval.argtypes = [C.c_void_p, C.POINTER(MLIS_Tensor), C.POINTER(MLIS_Tensor), C.POINTER(MLIS_Tensor), C.c_float, C.POINTER(MLIS_Tensor)]
    return L


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(HOST_LIB) or not os.path.exists(ENGINE_LIB):
            raise EngineMissing("%s / %s not built (run __graft_entry__.build()); there is no CPU fallback" % (HOST_LIB, ENGINE_LIB))
        _lib = bind(C.CDLL(HOST_LIB, mode=C.RTLD_LOCAL))
    return _lib


class MLISError(RuntimeError):
    pass


def _to_tensor(arr):
    """numpy (C-order, reversed ggml shape) -> MLIS_Tensor view (no copy; keep `arr` alive)."""
    arr = np.ascontiguousarray(arr, dtype=np.float32)
    ne = list(arr.shape[::-1]) + [1] * (4 - arr.ndim)
    t = MLIS_Tensor(arr.ctypes.data_as(C.POINTER(C.c_float)), (C.c_int * 4)(*ne), 0)
    t._keep = arr
    return t


def _from_tensor(t):
    n = [int(x) for x in t.n]
    cnt = n[0] * n[1] * n[2] * n[3]
    return np.ctypeslib.as_array(t.d, shape=(cnt,)).reshape(n[::-1]).copy()


class Ctx:
    def __init__(self, _lib=None, **opts):
        self.L = _lib if _lib is not None else lib()
        self.h = C.c_void_p(self.L.mlis_ctx_create_i(0x000402))
        self._cb = None
        for k, v in opts.items():
            self.set(k, v)

    def close(self):
        if self.h:
            self.L.mlis_ctx_destroy(C.byref(self.h))
            self.h = None

    def _chk(self, r):
        if r < 0:
            raise MLISError("%d: %s" % (r, self.L.mlis_errstr_get(self.h).decode(errors="replace")))
        return r

    def set(self, name, value):
        if isinstance(value, (tuple, list)):
            value = ",".join(str(v) for v in value)
        if isinstance(value, bool):
            value = int(value)
        return self._chk(self.L.mlis_option_set_str(self.h, name.encode(), str(value).encode()))

    def set_image(self, img_u8, mask=False):
        img_u8 = np.ascontiguousarray(img_u8, dtype=np.uint8)
        h, w = img_u8.shape[:2]
        c = 1 if img_u8.ndim == 2 else img_u8.shape[2]
        im = MLIS_Image(img_u8.ctypes.data_as(C.POINTER(C.c_uint8)), img_u8.size, w, h, c, 0)
        f = self.L.mlis_option_set
        f.argtypes = [C.c_void_p, C.c_int, C.POINTER(MLIS_Image)]
        return self._chk(f(self.h, OPT_IMAGE_MASK if mask else OPT_IMAGE, C.byref(im)))

    def set_callback(self, fn):
        self._cb = CALLBACK(lambda u, c, p: fn(p.contents) or 0)
        f = self.L.mlis_option_set
        f.argtypes = [C.c_void_p, C.c_int, CALLBACK, C.c_void_p]
        return self._chk(f(self.h, OPT_CALLBACK, self._cb, None))

    def setup(self):
        return self._chk(self.L.mlis_setup(self.h))

    def generate(self):
        return self._chk(self.L.mlis_generate(self.h))

    def image(self, idx=0):
        p = self.L.mlis_image_get(self.h, idx)
        if not p:
            raise MLISError(self.L.mlis_errstr_get(self.h).decode())
        im = p.contents
        return np.ctypeslib.as_array(im.d, shape=(im.h, im.w, im.c)).copy()

    def tensor(self, tid):
        p = self.L.mlis_tensor_get(self.h, tid)
        if not p or not p.contents.d:
            return None
        return _from_tensor(p.contents)

    def tensor_set(self, tid, arr):
        p = self.L.mlis_tensor_get(self.h, tid)
        arr = np.ascontiguousarray(arr, dtype=np.float32)
        ne = list(arr.shape[::-1]) + [1] * (4 - arr.ndim)
        self.L.mlis_tensor_resize(p, *ne)
        C.memmove(p.contents.d, arr.ctypes.data, arr.nbytes)

    def infotext(self):
        s = self.L.mlis_infotext_get(self.h, 0)
        return s.decode() if s else ""

    def tokenize(self, text, model=SUBMODEL_CLIP):
        pt = C.POINTER(C.c_int32)()
        n = self._chk(self.L.mlis_text_tokenize(self.h, text.encode(), C.byref(pt), model))
        return [int(pt[i]) for i in range(n)]

    def clip_encode(self, text, model=SUBMODEL_CLIP, feat=False, flags=0, both=False):
        """embed [77,d] (feat=False), pooled projected feature [d] (feat=True), or (embed, feat) from one call (both=True)."""
        e, f = MLIS_Tensor(), MLIS_Tensor()
        want_e, want_f = both or not feat, both or feat
        self._chk(self.L.mlis_clip_text_encode(self.h, text.encode(), C.byref(e) if want_e else None, C.byref(f) if want_f else None, model, flags))
        oe = _from_tensor(e) if want_e else None
        of = _from_tensor(f) if want_f else None
        self.L.mlis_tensor_free(C.byref(e)); self.L.mlis_tensor_free(C.byref(f))
        return (oe, of) if both else (of if feat else oe)

    def unet_eval(self, x, cond, label, sigma):
        tx, tc = _to_tensor(x), _to_tensor(cond)
        tl = _to_tensor(label) if label is not None else None
        dx = MLIS_Tensor()
        self._chk(self.L.mlis_unet_eval(self.h, C.byref(tx), C.byref(tc), C.byref(tl) if tl is not None else None, C.c_float(sigma), C.byref(dx)))
        out = _from_tensor(dx)
        self.L.mlis_tensor_free(C.byref(dx))
        return out

    def decode(self, latent):
        tl = _to_tensor(latent)
        img = MLIS_Tensor()
        self._chk(self.L.mlis_image_decode(self.h, C.byref(tl), C.byref(img), 0))
        out = _from_tensor(img)
        self.L.mlis_tensor_free(C.byref(img))
        return out

    def cfg_split(self, half, exchange=None):
        """Cross-GPU CFG split: this context evaluates CFG half `half` (0 cond, 1 uncond); exchange(mine_ptr, other_ptr, n)
        must fill the peer's output (device pointers, n floats). half=None switches it off."""
        if half is None:
            self._cfg_cb = None
            return self._chk(self.L.mlis_b200_cfg_split_set(self.h, -1, CFG_EXCHANGE(0), None))
        def cb(user, mine, other, n):
            try:
                exchange(mine, other, n)
                return 1
            except Exception as e:      # never let an exception cross the C boundary
                print("cfg exchange failed:", e)
                return -1
        self._cfg_cb = CFG_EXCHANGE(cb)
        return self._chk(self.L.mlis_b200_cfg_split_set(self.h, half, self._cfg_cb, None))

    def images_device(self):
        """(device pointer, n, h, w) of the RGB8 images of the last generation as they lie in HBM."""
        p, w, h, n = C.c_void_p(), C.c_int(), C.c_int(), C.c_int()
        if self.L.mlis_b200_images_device(self.h, C.byref(p), C.byref(w), C.byref(h), C.byref(n)) < 0:
            raise MLISError(self.L.mlis_errstr_get(self.h).decode())
        return p.value, n.value, h.value, w.value

    # ---- VAE decode tiles spread across GPUs (include/mlimgsynth_b200.h)
    def vae_tile_plan(self, lw, lh):
        """(n_tiles, tile_w_px, tile_h_px) of the current `vae_tile` option for a latent of lw x lh."""
        n, tw, th = C.c_int(), C.c_int(), C.c_int()
        self._chk(self.L.mlis_b200_vae_tile_plan(self.h, lw, lh, C.byref(n), C.byref(tw), C.byref(th)))
        return n.value, tw.value, th.value

    def vae_tiles_decode(self, latent, rank, world, tiles_dev_ptr):
        """Decode tiles rank, rank + world, ... of `latent` [1,4,lh,lw] into the caller's device buffer."""
        tl = _to_tensor(latent)
        self._chk(self.L.mlis_b200_vae_tiles_decode(self.h, C.byref(tl), rank, world, C.c_void_p(tiles_dev_ptr)))

    def vae_tiles_merge(self, lw, lh, gathered_dev_ptr, world, slots_per_worker, want_float=False):
        """Paste the gathered tiles in the reference's order; the image is then available through image(0)."""
        img = MLIS_Tensor() if want_float else None
        self._chk(self.L.mlis_b200_vae_tiles_merge(self.h, lw, lh, C.c_void_p(gathered_dev_ptr), world, slots_per_worker, C.byref(img) if want_float else None))
        if want_float:
            out = _from_tensor(img)
            self.L.mlis_tensor_free(C.byref(img))
            return out

    def encode(self, image):
        ti = _to_tensor(image)
        lat = MLIS_Tensor()
        self._chk(self.L.mlis_image_encode(self.h, C.byref(ti), C.byref(lat), 0))
        out = _from_tensor(lat)
        self.L.mlis_tensor_free(C.byref(lat))
        return out
