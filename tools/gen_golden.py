#!/usr/bin/env python3
"""Generates tests/golden/host_kat.json from the REFERENCE's own compiled host code
(oracle/_ref/libmlimgsynth_cpu.so, built from /root/reference by oracle/Makefile). The fixtures pin
the bit-exact host pieces of the path: Philox noise stream, CLIP tokenizer, sigma <-> t maps.
Run in the build container (the reference tree is not on the GPU box); the JSON is committed."""
import ctypes as C, json, os, struct
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
L = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libmlimgsynth_cpu.so"))

class Rng(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("offset", C.c_uint32)]

out = {}
# --- RNG: (seed, offset, n) -> float bit patterns
rng = []
for seed, offset, n in [(0, 0, 12), (42, 0, 16), (42, 7, 5), (2**40 + 5, 3, 9), (12345678901234567, 1000, 8)]:
    r = Rng(seed, offset); buf = (C.c_float * n)()
    L.rng_philox_randn(C.byref(r), n, buf)
    rng.append({"seed": seed, "offset": offset, "n": n, "bits": [struct.unpack("<I", struct.pack("<f", x))[0] for x in buf], "offset_after": r.offset})
out["rng"] = rng
# --- tokenizer via the public API (no model needed when the type is forced, test_text_tokenize_clip.c:38)
L.mlis_ctx_create_i.restype = C.c_void_p
ctx = C.c_void_p(L.mlis_ctx_create_i(0x000402))
L.mlis_option_set_str.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
assert L.mlis_option_set_str(ctx, b"model_type", b"sd1") >= 0
L.mlis_text_tokenize.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.POINTER(C.c_int32)), C.c_int]
texts = ["a dog jumping", "   a   dog\t\tjumping\r\n", "an illustration", "a sign saying \"Here lies Cesar\"", "a sign saying 'Here lies Cesar'",
         "2025", "A'veA'llA's", "", "  \t  \n", "a dog, a house.", "corazón", "cat---dog-—-rabbit",
         "まあ、お待ちなさい。",
         "Stable Diffusion is a deep learning, text-to-image model released in 2022 based on diffusion techniques.",
         "I'd say it's 3.14% OK; don't YOU'RE we'VE", "ÉCOLE Über STRASSE İstanbul Σοφία", "emoji \U0001F600 test nbsp　ideographic",
         "masterpiece, best quality, (1girl:1.2), highres, 8k, <lora:x:0.5>", "x" * 70, "hello_world-foo/bar\\baz 1st 22nd 3.5e-7"]
tok = []
for t in texts:
    p = C.POINTER(C.c_int32)()
    n = L.mlis_text_tokenize(ctx, t.encode("utf-8"), C.byref(p), 4)
    tok.append({"text": t, "ids": [int(p[i]) for i in range(n)] if n >= 0 else None, "ret": n})
out["tokenizer"] = tok
# --- sigma tables
L.unet_params_init()
L.unet_sigma_to_t.restype = C.c_float; L.unet_t_to_sigma.restype = C.c_float
L.unet_sigma_to_t.argtypes = [C.c_void_p, C.c_float]; L.unet_t_to_sigma.argtypes = [C.c_void_p, C.c_float]
P = C.addressof(C.c_char.in_dll(L, "g_unet_sd1"))
f2b = lambda x: struct.unpack("<I", struct.pack("<f", x))[0]
out["t_to_sigma"] = [{"t": t, "bits": f2b(L.unet_t_to_sigma(P, t))} for t in [0.0, 0.5, 1.0, 52.578947, 499.5, 699.3, 998.0, 999.0, 1200.0, -3.0]]
out["sigma_to_t"] = [{"sigma": s, "bits": f2b(L.unet_sigma_to_t(P, s))} for s in [0.0291, 0.029167158, 0.05, 0.5, 1.0, 3.3251, 7.0, 14.0, 14.614641, 20.0]]
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "host_kat.json"), "w"), indent=0, ensure_ascii=True)
print({k: len(v) for k, v in out.items()})
