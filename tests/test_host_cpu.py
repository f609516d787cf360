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
