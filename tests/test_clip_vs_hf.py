"""End-to-end anchor for the text-conditioning path and, through it, for the oracle's arithmetic:
the reference's OWN host code (clip.c + mlblock: oracle/_ref/libmlimgsynth_cpu.so, compiled unmodified from
/root/reference) running on the CPU oracle must reproduce Hugging Face's CLIPTextModel -- the canonical implementation
of the SD1.x text encoder -- on the same random-init ViT-L/14 weights and token ids. This exercises token + position
embedding, 12 pre-LN blocks with causal attention and quick-GELU, and the final LayerNorm; the agreement (5e-4 of the
largest activation) is what f16 rounding of the linear operands leaves.

CPU only. The weights are the CLIP part of tools/gen_weights.py's SD1 checkpoint (246 MB, generated in a temp dir)."""
import ctypes as C
import json, os, struct, sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
PROMPT = b"a photograph of an astronaut riding a horse"
SUBMODEL_CLIP = 4


def read_safetensors(path, prefix):
    import torch
    out = {}
    with open(path, "rb") as f:
        n = struct.unpack("<Q", f.read(8))[0]
        hdr = json.loads(f.read(n)); base = 8 + n
        for k, v in hdr.items():
            if not k.startswith(prefix):
                continue
            f.seek(base + v["data_offsets"][0])
            raw = f.read(v["data_offsets"][1] - v["data_offsets"][0])
            out[k[len(prefix):]] = torch.from_numpy(np.frombuffer(raw, dtype=np.float16).reshape(v["shape"]).astype(np.float32))
    return out


def openclip_to_hf(sd, d):
    """OpenCLIP text-tower names (SD2.x cond_stage_model.model.*, fused in_proj) -> Hugging Face CLIPTextModel names."""
    out = {"text_model.embeddings.token_embedding.weight": sd["token_embedding.weight"],
           "text_model.embeddings.position_embedding.weight": sd["positional_embedding"],
           "text_model.final_layer_norm.weight": sd["ln_final.weight"], "text_model.final_layer_norm.bias": sd["ln_final.bias"]}
    ren = {"ln_1": "layer_norm1", "ln_2": "layer_norm2", "attn.out_proj": "self_attn.out_proj", "mlp.c_fc": "mlp.fc1", "mlp.c_proj": "mlp.fc2"}
    for k, v in sd.items():
        if not k.startswith("transformer.resblocks."):
            continue
        l, rest = k[len("transformer.resblocks."):].split(".", 1)
        q = "text_model.encoder.layers.%s." % l
        if rest.startswith("attn.in_proj_"):
            kind = rest[len("attn.in_proj_"):]
            for i, n in enumerate(("q_proj", "k_proj", "v_proj")):
                out[q + "self_attn.%s.%s" % (n, kind)] = v[i * d:(i + 1) * d].clone()
            continue
        for a, b in ren.items():
            if rest.startswith(a + "."):
                out[q + b + rest[len(a):]] = v
    return out


def test_reference_openclip_on_oracle_matches_hf(oracle_built, tmp_path):
    """SD2.x text encoder (OpenCLIP ViT-H/14): fused in_proj split into q/k/v on load (mlimgsynth.c:990-1030), 23 of the 24
    blocks (clip_skip 2, mlimgsynth.c:766), then ln_final; padding token 0. ggml_gelu is the tanh form, so the Hugging Face
    side uses gelu_pytorch_tanh."""
    torch = pytest.importorskip("torch")
    tr = pytest.importorskip("transformers")
    from mlimgsynth_b200.api import MLIS_Tensor
    import gen_weights
    lib_path = os.path.join(oracle_built, "libmlimgsynth_cpu.so")
    if not os.path.exists(lib_path):
        pytest.skip("oracle/_ref/libmlimgsynth_cpu.so not built")
    wpath = str(tmp_path / "clip_h14.safetensors")
    gen_weights.write_safetensors(wpath, gen_weights.build_spec("sd2", parts=("clip",)), 1234, "f16")
    L = C.CDLL(lib_path, mode=C.RTLD_LOCAL)
    L.mlis_ctx_create_i.restype = C.c_void_p; L.mlis_ctx_create_i.argtypes = [C.c_int]
    L.mlis_ctx_destroy.argtypes = [C.POINTER(C.c_void_p)]
    L.mlis_option_set_str.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
    L.mlis_errstr_get.restype = C.c_char_p; L.mlis_errstr_get.argtypes = [C.c_void_p]
    L.mlis_clip_text_encode.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(MLIS_Tensor), C.POINTER(MLIS_Tensor), C.c_int, C.c_int]
    L.mlis_text_tokenize.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.POINTER(C.c_int32)), C.c_int]
    h = C.c_void_p(L.mlis_ctx_create_i(0x000402))
    for k, v in (("backend", "CPU"), ("model", wpath), ("model_type", "sd2")):
        assert L.mlis_option_set_str(h, k.encode(), v.encode()) >= 0, L.mlis_errstr_get(h)
    e = MLIS_Tensor()
    assert L.mlis_clip_text_encode(h, PROMPT, C.byref(e), None, SUBMODEL_CLIP, 0) >= 1, L.mlis_errstr_get(h)
    assert [int(x) for x in e.n][:2] == [1024, 77]
    ref = np.ctypeslib.as_array(e.d, shape=(77 * 1024,)).reshape(77, 1024).copy()
    pt = C.POINTER(C.c_int32)()
    nt = L.mlis_text_tokenize(h, PROMPT, C.byref(pt), SUBMODEL_CLIP)
    toks = [int(pt[i]) for i in range(nt)]
    L.mlis_ctx_destroy(C.byref(h))

    cfg = tr.CLIPTextConfig(vocab_size=49408, hidden_size=1024, intermediate_size=4096, num_hidden_layers=24, num_attention_heads=16,
                            max_position_embeddings=77, hidden_act="gelu_pytorch_tanh")
    m = tr.CLIPTextModel(cfg).eval()
    missing, unexpected = m.load_state_dict(openclip_to_hf(read_safetensors(wpath, "cond_stage_model.model."), 1024), strict=False)
    assert not missing and not unexpected
    ids = [49406] + toks + [49407]
    ids += [0] * (77 - len(ids))
    with torch.no_grad():
        o = m(input_ids=torch.tensor([ids]), output_hidden_states=True)
        want = m.text_model.final_layer_norm(o.hidden_states[-2])[0].numpy()
    err = np.abs(ref - want).max() / np.abs(want).max()
    print("reference OpenCLIP on the oracle vs HF: max-rel err %.2e" % err)
    assert err <= 3e-3


def test_reference_clip_on_oracle_matches_hf_clip_text_model(oracle_built, tmp_path):
    torch = pytest.importorskip("torch")
    tr = pytest.importorskip("transformers")
    from mlimgsynth_b200.api import MLIS_Tensor
    import gen_weights
    lib_path = os.path.join(oracle_built, "libmlimgsynth_cpu.so")
    if not os.path.exists(lib_path):
        pytest.skip("oracle/_ref/libmlimgsynth_cpu.so not built")
    wpath = str(tmp_path / "clip_l14.safetensors")
    gen_weights.write_safetensors(wpath, gen_weights.build_spec("sd1", parts=("clip",)), 1234, "f16")

    # ---- the reference's host code on the oracle
    L = C.CDLL(lib_path, mode=C.RTLD_LOCAL)
    L.mlis_ctx_create_i.restype = C.c_void_p; L.mlis_ctx_create_i.argtypes = [C.c_int]
    L.mlis_ctx_destroy.argtypes = [C.POINTER(C.c_void_p)]
    L.mlis_option_set_str.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
    L.mlis_errstr_get.restype = C.c_char_p; L.mlis_errstr_get.argtypes = [C.c_void_p]
    L.mlis_clip_text_encode.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(MLIS_Tensor), C.POINTER(MLIS_Tensor), C.c_int, C.c_int]
    L.mlis_text_tokenize.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.POINTER(C.c_int32)), C.c_int]
    h = C.c_void_p(L.mlis_ctx_create_i(0x000402))
    for k, v in (("backend", "CPU"), ("model", wpath), ("model_type", "sd1")):
        assert L.mlis_option_set_str(h, k.encode(), v.encode()) >= 0, L.mlis_errstr_get(h)
    e = MLIS_Tensor()
    assert L.mlis_clip_text_encode(h, PROMPT, C.byref(e), None, SUBMODEL_CLIP, 0) >= 1, L.mlis_errstr_get(h)
    n = [int(x) for x in e.n]
    assert n[:2] == [768, 77]
    ref = np.ctypeslib.as_array(e.d, shape=(77 * 768,)).reshape(77, 768).copy()
    pt = C.POINTER(C.c_int32)()
    nt = L.mlis_text_tokenize(h, PROMPT, C.byref(pt), SUBMODEL_CLIP)
    toks = [int(pt[i]) for i in range(nt)]
    assert toks == [320, 8853, 539, 550, 18376, 6765, 320, 4558]        # ids of the reference's own tokenizer tests' vocabulary
    L.mlis_ctx_destroy(C.byref(h))

    # ---- Hugging Face CLIPTextModel, same weights, same ids (BOS + tokens + EOS, padded with EOS: clip.c:23-35)
    cfg = tr.CLIPTextConfig(vocab_size=49408, hidden_size=768, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=12,
                            max_position_embeddings=77, hidden_act="quick_gelu")
    m = tr.CLIPTextModel(cfg).eval()
    missing, unexpected = m.load_state_dict(read_safetensors(wpath, "cond_stage_model.transformer."), strict=False)
    assert not missing and not unexpected
    ids = [49406] + toks + [49407]
    ids += [49407] * (77 - len(ids))
    with torch.no_grad():
        want = m(input_ids=torch.tensor([ids])).last_hidden_state[0].numpy()
    err = np.abs(ref - want).max() / np.abs(want).max()
    print("reference CLIP on the oracle vs HF CLIPTextModel: max-rel err %.2e" % err)
    assert err <= 2e-3


def test_reference_sdxl_second_encoder_and_pooled_feature_match_hf(oracle_built, tmp_path):
    """SDXL's second text encoder (OpenCLIP bigG/14, 32 blocks): the penultimate hidden state without the final norm
    (what goes into the 2048-wide context, mlimgsynth.c:1502-1563) and the pooled feature -- ln_final of the last block at
    the first end-of-text token times text_projection (clip.c:418-437), the head of the label vector -- against
    CLIPTextModelWithProjection."""
    torch = pytest.importorskip("torch")
    tr = pytest.importorskip("transformers")
    from mlimgsynth_b200.api import MLIS_Tensor
    import gen_weights
    lib_path = os.path.join(oracle_built, "libmlimgsynth_cpu.so")
    if not os.path.exists(lib_path):
        pytest.skip("oracle/_ref/libmlimgsynth_cpu.so not built")
    wpath = str(tmp_path / "clip_sdxl.safetensors")
    gen_weights.write_safetensors(wpath, gen_weights.build_spec("sdxl", parts=("clip",)), 1234, "f16")
    L = C.CDLL(lib_path, mode=C.RTLD_LOCAL)
    L.mlis_ctx_create_i.restype = C.c_void_p; L.mlis_ctx_create_i.argtypes = [C.c_int]
    L.mlis_ctx_destroy.argtypes = [C.POINTER(C.c_void_p)]
    L.mlis_option_set_str.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
    L.mlis_errstr_get.restype = C.c_char_p; L.mlis_errstr_get.argtypes = [C.c_void_p]
    L.mlis_clip_text_encode.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(MLIS_Tensor), C.POINTER(MLIS_Tensor), C.c_int, C.c_int]
    L.mlis_text_tokenize.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.POINTER(C.c_int32)), C.c_int]
    SUBMODEL_CLIP2, NO_NORM = 5, 1
    h = C.c_void_p(L.mlis_ctx_create_i(0x000402))
    for k, v in (("backend", "CPU"), ("model", wpath), ("model_type", "sdxl")):
        assert L.mlis_option_set_str(h, k.encode(), v.encode()) >= 0, L.mlis_errstr_get(h)
    e, f = MLIS_Tensor(), MLIS_Tensor()
    assert L.mlis_clip_text_encode(h, PROMPT, C.byref(e), None, SUBMODEL_CLIP2, NO_NORM) >= 1, L.mlis_errstr_get(h)
    assert L.mlis_clip_text_encode(h, PROMPT, None, C.byref(f), SUBMODEL_CLIP2, 0) >= 1, L.mlis_errstr_get(h)
    assert [int(x) for x in e.n][:2] == [1280, 77] and int(f.n[0]) == 1280
    emb = np.ctypeslib.as_array(e.d, shape=(77 * 1280,)).reshape(77, 1280).copy()
    feat = np.ctypeslib.as_array(f.d, shape=(1280,)).copy()
    pt = C.POINTER(C.c_int32)()
    nt = L.mlis_text_tokenize(h, PROMPT, C.byref(pt), SUBMODEL_CLIP2)
    toks = [int(pt[i]) for i in range(nt)]
    L.mlis_ctx_destroy(C.byref(h))

    sd = read_safetensors(wpath, "conditioner.embedders.1.model.")
    hf = openclip_to_hf(sd, 1280)
    hf["text_projection.weight"] = sd["text_projection"].t().contiguous()       # OpenCLIP applies x @ P, nn.Linear x @ W^T
    cfg = tr.CLIPTextConfig(vocab_size=49408, hidden_size=1280, intermediate_size=5120, num_hidden_layers=32, num_attention_heads=20,
                            max_position_embeddings=77, hidden_act="gelu_pytorch_tanh", projection_dim=1280,
                            eos_token_id=49407, pad_token_id=0, bos_token_id=49406)
    m = tr.CLIPTextModelWithProjection(cfg).eval()
    missing, unexpected = m.load_state_dict(hf, strict=False)
    assert not missing and not unexpected
    ids = [49406] + toks + [49407]
    ids += [0] * (77 - len(ids))
    with torch.no_grad():
        o = m(input_ids=torch.tensor([ids]), output_hidden_states=True)
    want_emb, want_feat = o.hidden_states[-2][0].numpy(), o.text_embeds[0].numpy()
    e1 = np.abs(emb - want_emb).max() / np.abs(want_emb).max()
    e2 = np.abs(feat - want_feat).max() / np.abs(want_feat).max()
    print("SDXL encoder 2 on the oracle vs HF: hidden %.2e, pooled feature %.2e" % (e1, e2))
    assert e1 <= 3e-3 and e2 <= 3e-3
