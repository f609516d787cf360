"""Bit-exact host pieces of the path (no GPU): Philox noise stream, CLIP tokenizer, prompt-emphasis
parser, sigma<->t maps of the B200 host layer (libmlimgsynth_b200.so) against
  (1) the golden vectors generated from the reference's own compiled code (tests/golden/host_kat.json,
      tools/gen_golden.py), which include the reference's known-answer tests (test_rng.c:11-24,
      test_text_tokenize_clip.c:41-66), and
  (2) the reference's prompt-parser test cases (test_prompt_preproc.c:101-127), restated here.
"""
import ctypes as C, json, os, struct, sys
import pytest
import mlimgsynth_b200
from mlimgsynth_b200 import api

GOLD = json.load(open(os.path.join(mlimgsynth_b200.ROOT, "tests", "golden", "host_kat.json")))
f2b = lambda x: struct.unpack("<I", struct.pack("<f", x))[0]


class Rng(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("offset", C.c_uint32)]


def test_rng_reference_kat_present():
    # test_rng.c:11-24 (seed 0, offset 0): first values -0.92466259 -0.42534414 -2.64384580
    g = GOLD["rng"][0]
    vals = [struct.unpack("<f", struct.pack("<I", b))[0] for b in g["bits"]]
    assert ["%.8f" % v for v in vals[:3]] == ["-0.92466259", "-0.42534414", "-2.64384580"]


@pytest.mark.parametrize("case", GOLD["rng"])
def test_rng_bit_exact(case):
    L = api.lib()
    r = Rng(case["seed"], case["offset"]); buf = (C.c_float * case["n"])()
    L.rng_philox_randn(C.byref(r), case["n"], buf)
    assert [f2b(x) for x in buf] == case["bits"]
    assert r.offset == case["offset_after"]


@pytest.fixture(scope="module")
def tok_ctx():
    c = api.Ctx(model_type="sd1")
    yield c
    c.close()


@pytest.mark.parametrize("case", GOLD["tokenizer"], ids=lambda c: repr(c["text"][:24]))
def test_tokenizer_bit_exact(tok_ctx, case):
    assert tok_ctx.tokenize(case["text"]) == case["ids"]


def test_tokenizer_reference_kats(tok_ctx):
    # test_text_tokenize_clip.c:41-66
    assert tok_ctx.tokenize("a dog jumping") == [320, 1929, 11476]
    assert tok_ctx.tokenize("2025") == [17, 15, 17, 276]
    assert tok_ctx.tokenize("A'veA'llA's") == [320, 1200, 320, 1342, 320, 568]
    assert tok_ctx.tokenize("cat---dog-—-rabbit") == [2368, 11079, 1929, 12, 6718, 268, 10274]
    assert tok_ctx.tokenize("") == []


@pytest.mark.parametrize("key,fn", [("t_to_sigma", "unet_t_to_sigma"), ("sigma_to_t", "unet_sigma_to_t")])
def test_sigma_maps_bit_exact(key, fn):
    L = api.lib()
    L.unet_params_init()
    f = getattr(L, fn); f.restype = C.c_float; f.argtypes = [C.c_void_p, C.c_float]
    P = C.addressof(C.c_char.in_dll(L, "g_unet_sd1"))
    for c in GOLD[key]:
        x = c["t"] if key == "t_to_sigma" else c["sigma"]
        assert f2b(f(P, x)) == c["bits"], (x, f(P, x))


# ---- prompt parser: test_prompt_preproc.c:101-127 restated
class Chunk(C.Structure):
    _fields_ = [("beg", C.c_int), ("len", C.c_int), ("w", C.c_float)]


class PT(C.Structure):
    _fields_ = [("text", C.c_char_p), ("text_len", C.c_int), ("text_cap", C.c_int), ("data", C.POINTER(C.c_char)), ("data_len", C.c_int),
                ("data_cap", C.c_int), ("chunks", C.POINTER(Chunk)), ("n_chunks", C.c_int), ("cap_chunks", C.c_int),
                ("loras", C.POINTER(Chunk)), ("n_loras", C.c_int), ("cap_loras", C.c_int)]


def parse(text, raw=False):
    L = api.lib()
    p = PT()
    b = text.encode()
    if raw:
        L.prompt_text_set_raw(C.byref(p), b, len(b)); r = 1
    else:
        r = L.prompt_text_set_parse(C.byref(p), b, len(b))
    if r < 0:
        L.prompt_text_free(C.byref(p)); return r, None, None
    t = p.text[:p.text_len] if p.text else b""
    chunks = [(t[p.chunks[i].beg:p.chunks[i].beg + p.chunks[i].len].decode(), p.chunks[i].w) for i in range(p.n_chunks)]
    loras = [(C.string_at(C.addressof(p.data.contents) + p.loras[i].beg, p.loras[i].len).decode(), p.loras[i].w) for i in range(p.n_loras)]
    L.prompt_text_free(C.byref(p))
    return r, chunks, loras


F = lambda x: struct.unpack("<f", struct.pack("<f", x))[0]


@pytest.mark.parametrize("text,chunks,loras", [
    ("a b c", [("a b c", 1.0)], []),
    ("a (b) c", [("a ", 1.0), ("b", F(1.1)), (" c", 1.0)], []),
    ("a ((b)) c", [("a ", 1.0), ("b", F(1.1 ** 2)), (" c", 1.0)], []),
    ("a [b] c", [("a ", 1.0), ("b", F(1.1 ** -1)), (" c", 1.0)], []),
    ("a (b:1.5) c", [("a ", 1.0), ("b", 1.5), (" c", 1.0)], []),
    ("a \\(b\\) c", [("a (b) c", 1.0)], []),
    ("a <lora:NAME:0.8> c", [("a  c", 1.0)], [("NAME", F(0.8))]),
    ("a <lora:NAME> c", [("a  c", 1.0)], [("NAME", 1.0)]),
    ("a BREAK c", [("a  c", 1.0)], []),
    ("(a) b", [("a", F(1.1)), (" b", 1.0)], []),
])
def test_prompt_parse(text, chunks, loras):
    r, c, l = parse(text)
    assert r > 0
    assert [(t, f2b(w)) for t, w in c] == [(t, f2b(w)) for t, w in chunks]
    assert [(t, f2b(w)) for t, w in l] == [(t, f2b(w)) for t, w in loras]


def test_prompt_parse_errors_and_raw():
    assert parse("a ) b")[0] == -5
    assert parse("a [b:1.5] c")[0] == -5
    assert parse("a <lora:x")[0] == -5
    assert parse("a <foo:x> b")[0] == -5
    r, c, l = parse("a (b:1.5) <lora:x>", raw=True)
    assert c == [("a (b:1.5) <lora:x>", 1.0)] and l == []


def test_option_api_errors():
    c = api.Ctx()
    with pytest.raises(api.MLISError):
        c.set("no_such_option", 1)
    with pytest.raises(api.MLISError):
        c.set("steps", "abc")
    c.set("method", "dpm++2m"); c.set("scheduler", "karras"); c.set("image-dim", (512, 768))
    with pytest.raises(api.MLISError):   # no model set: fails loudly, no CPU fallback
        c.generate()
    c.close()


def test_tensor_name_conversion_matches_reference(oracle_built):
    """Weight ingestion: every checkpoint key of the SD1.x / SD2.x / SDXL architectures (CLIP HF and OpenCLIP
    towers, VAE, UNet; 3276 names) plus keys the loader must ignore map through host/name_conv.c exactly as through the
    reference's tensor_name_conv.c:274 (compiled unmodified into oracle/_ref): same result code (unused / good / fused-QKV)
    and the same internal name."""
    import ctypes as C
    sys.path.insert(0, os.path.join(mlimgsynth_b200.ROOT, "tools"))
    import gen_weights

    class StrSlice(C.Structure):
        _fields_ = [("b", C.c_char_p), ("s", C.c_size_t)]
    ref_lib = os.path.join(oracle_built, "libmlimgsynth_cpu.so")
    if not os.path.exists(ref_lib):
        pytest.skip("oracle/_ref/libmlimgsynth_cpu.so not built")
    R = C.CDLL(ref_lib); R.tnconv_sd.argtypes = [StrSlice, C.POINTER(C.c_void_p)]
    O = C.CDLL(mlimgsynth_b200.HOST_LIB); O.tnconv_sd.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t]
    names = []
    for kind in ("sd1", "sd2", "sdxl"):
        names += [n for n, _, _ in gen_weights.build_spec(kind).items]
    names += ["model_ema.decay", "alphas_cumprod", "cond_stage_model.transformer.text_model.embeddings.position_ids",
              "first_stage_model.loss.logvar", "foo.bar", "model.diffusion_model.unknown_block.0.weight"]
    names = list(dict.fromkeys(names))
    assert len(names) > 3000
    for n in names:
        b = n.encode()
        out = C.c_void_p(None)
        r = R.tnconv_sd(StrSlice(b, len(b)), C.byref(out))
        want = C.string_at(out.value).decode() if out.value else ""
        buf = C.create_string_buffer(512)
        o = O.tnconv_sd(b, buf, 512)
        assert (r > 0) == (o > 0), (n, r, o)
        if r > 0:
            assert (o, buf.value.decode()) == (r, want), (n, r, want, o, buf.value.decode())


def test_tensor_store_reads_and_converts_safetensors(tmp_path):
    """Weight ingestion: safetensors index, shape reversal to ggml order (tensorstore_safet.c:138-142), name
    mapping through the converter, dropped unknown keys, and the dtype conversions the engine asks for: F32 -> F16 rounds to
    nearest even like ggml_fp32_to_fp16_row, BF16 -> F32 is exact, BF16 -> F16 goes through F32."""
    import numpy as np
    rng = np.random.default_rng(3)
    a32 = (rng.standard_normal((5, 7)) * 3).astype(np.float32)
    a32[0, :4] = [65504.0, 1e-8, -0.0, 6.1e-5]                       # f16 max, underflow, signed zero, smallest normal region
    a16 = rng.standard_normal((3, 2, 4)).astype(np.float16)
    bf_src = rng.standard_normal(11).astype(np.float32)
    abf = (bf_src.view(np.uint32) >> 16).astype(np.uint16)           # bf16 bit patterns (truncation is fine for a fixture)
    tensors = [("model.diffusion_model.time_embed.0.weight", "F32", a32.shape, a32.tobytes()),
               ("model.diffusion_model.time_embed.0.bias", "BF16", abf.shape, abf.tobytes()),
               ("first_stage_model.post_quant_conv.weight", "F16", a16.shape, a16.tobytes()),
               ("model_ema.decay", "F32", (1,), np.zeros(1, np.float32).tobytes())]
    header, off = {}, 0
    for name, dt, shape, raw in tensors:
        header[name] = {"dtype": dt, "shape": list(shape), "data_offsets": [off, off + len(raw)]}
        off += len(raw)
    hj = json.dumps(header, separators=(",", ":")).encode()
    hj += b" " * ((8 - len(hj) % 8) % 8)
    path = str(tmp_path / "mini.safetensors")
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(hj))); f.write(hj)
        for _, _, _, raw in tensors:
            f.write(raw)

    class TSEntry(C.Structure):
        _fields_ = [("key", C.c_char_p), ("dtype", C.c_int), ("ndim", C.c_int), ("shape", C.c_int64 * 4), ("data", C.c_void_p),
                    ("nbytes", C.c_size_t), ("owned", C.c_bool)]

    class TStore(C.Structure):
        _fields_ = [("e", C.POINTER(TSEntry)), ("n", C.c_int), ("cap", C.c_int), ("hash", C.POINTER(C.c_int)), ("hash_cap", C.c_int),
                    ("maps", C.c_void_p), ("n_map", C.c_int), ("cap_map", C.c_int)]
    L = C.CDLL(mlimgsynth_b200.HOST_LIB)
    L.tstore_read_safetensors.argtypes = [C.POINTER(TStore), C.c_char_p, C.c_void_p, C.c_char_p]
    L.tstore_find.restype = C.POINTER(TSEntry); L.tstore_find.argtypes = [C.POINTER(TStore), C.c_char_p]
    L.tsentry_as.restype = C.c_void_p; L.tsentry_as.argtypes = [C.POINTER(TSEntry), C.c_int, C.POINTER(C.c_void_p)]
    L.tsentry_count.restype = C.c_int64; L.tsentry_count.argtypes = [C.POINTER(TSEntry)]
    conv = C.cast(L.tnconv_sd, C.c_void_p)
    S = TStore()
    assert L.tstore_read_safetensors(C.byref(S), path.encode(), conv, None) >= 1
    assert S.n == 3                                                  # model_ema.decay is not a tensor of the path: dropped

    def fetch(key, want, npdt, count):
        e = L.tstore_find(C.byref(S), key)
        assert e, key
        assert L.tsentry_count(e) == count
        tofree = C.c_void_p(None)
        p = L.tsentry_as(e, want, C.byref(tofree))
        assert p
        return np.frombuffer(C.string_at(p, count * np.dtype(npdt).itemsize), dtype=npdt).copy(), e.contents

    # names and ggml-order shapes
    k_w = [k for k in (b"unet.embed.0.weight", b"unet.time_embed.0.weight") if L.tstore_find(C.byref(S), k)]
    assert k_w, "time_embed weight not mapped"
    w16, ew = fetch(k_w[0], 1, np.float16, 35)
    assert list(ew.shape)[:2] == [7, 5] and ew.dtype == 0
    assert np.array_equal(w16.view(np.uint16), a32.astype(np.float16).reshape(-1).view(np.uint16))       # round to nearest even, bit for bit
    w32, _ = fetch(k_w[0], 0, np.float32, 35)
    assert np.array_equal(w32, a32.reshape(-1))
    k_b = k_w[0].replace(b"weight", b"bias")
    b32, eb = fetch(k_b, 0, np.float32, 11)
    assert eb.dtype == 2 and np.array_equal(b32.view(np.uint32), abf.astype(np.uint32) << 16)
    b16, _ = fetch(k_b, 1, np.float16, 11)
    assert np.array_equal(b16.view(np.uint16), (abf.astype(np.uint32) << 16).view(np.float32).astype(np.float16).view(np.uint16))
    k_v = [k for k in (b"vae.post_quant_conv.weight",) if L.tstore_find(C.byref(S), k)]
    assert k_v
    v16, ev = fetch(k_v[0], 1, np.float16, 24)
    assert list(ev.shape)[:3] == [4, 2, 3] and np.array_equal(v16.view(np.uint16), a16.reshape(-1).view(np.uint16))
    # a second file read into the same store (model + TAE) keeps BOTH mappings alive until tstore_free
    assert L.tstore_read_safetensors(C.byref(S), path.encode(), None, b"second.") >= 1 and S.n_map == 2
    w2, _ = fetch(k_w[0], 0, np.float32, 35)
    assert np.array_equal(w2, a32.reshape(-1))
    L.tstore_free(C.byref(S))
    assert S.n == 0 and S.n_map == 0

    # the header is not trusted: entries whose byte range is negative, outside the file or not shape x dtype are rejected
    def write(name, header_obj, payload, hlen=None):
        hj = json.dumps(header_obj, separators=(",", ":")).encode()
        q = str(tmp_path / name)
        with open(q, "wb") as f:
            f.write(struct.pack("<Q", len(hj) if hlen is None else hlen)); f.write(hj); f.write(payload)
        return q
    key = "model.diffusion_model.time_embed.0.weight"
    bad = [
        write("neg.safetensors", {key: {"dtype": "F32", "shape": [2, 2], "data_offsets": [-16, 0]}}, b"\0" * 16),
        write("short.safetensors", {key: {"dtype": "F32", "shape": [4, 4], "data_offsets": [0, 16]}}, b"\0" * 16),      # 16 bytes for 64
        write("past.safetensors", {key: {"dtype": "F16", "shape": [8], "data_offsets": [8, 24]}}, b"\0" * 16),          # beyond the data section
        write("rev.safetensors", {key: {"dtype": "F16", "shape": [8], "data_offsets": [16, 0]}}, b"\0" * 16),
        write("hdr.safetensors", {key: {"dtype": "F16", "shape": [8], "data_offsets": [0, 16]}}, b"\0" * 16, hlen=1 << 40),
    ]
    tiny = str(tmp_path / "tiny.safetensors"); open(tiny, "wb").write(b"abc")
    for q in bad + [tiny]:
        S2 = TStore()
        assert L.tstore_read_safetensors(C.byref(S2), q.encode(), conv, None) < 0, q
        assert S2.n == 0 and S2.n_map == 0, q                         # nothing indexed, mapping released
        L.tstore_free(C.byref(S2))
