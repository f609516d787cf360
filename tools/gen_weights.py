#!/usr/bin/env python3
"""Random-init checkpoint generator (SD1.x / SD2.x / SDXL / TAESD), CompVis/LDM tensor names.

There is no network, so the benchmark and the parity tests run on locally generated
random-init weights of the named architectures (BASELINE.json north_star). The file layout
is the one the reference loader expects: safetensors, LDM key names (renamed on load by
tensor_name_conv.c:274 `tnconv_sd`), PyTorch shape order (reversed to ggml order by
tensorstore_safet.c:138-142). Architectures follow unet.c:21-80, vae.c:21-44, clip.c:23-57,
tae.c:17-22.

Every tensor is drawn from its own stream seeded by crc32(name) ^ seed, so any subset can be
regenerated bit-identically (tests generate single blocks).

Initialisation is fp16-safe and roughly variance preserving (SURVEY.md section 7 "hard parts"):
  conv / linear weight  N(0, (gain/sqrt(fan_in))^2), bias N(0, 0.02^2)
  norm weight           1 + N(0, 0.05^2), bias N(0, 0.05^2)
  embeddings            N(0, 0.02^2) (token) / N(0, 0.01^2) (position)
"""
import argparse, json, struct, sys, zlib
import numpy as np

GAIN = 0.7


class Spec:
    """Ordered list of (name, shape, kind)."""
    def __init__(self):
        self.items = []

    def add(self, name, shape, kind):
        self.items.append((name, tuple(int(s) for s in shape), kind))

    def linear(self, name, n_in, n_out, bias=True):
        self.add(name + ".weight", (n_out, n_in), "w")
        if bias:
            self.add(name + ".bias", (n_out,), "b")

    def conv(self, name, c_in, c_out, k, bias=True):
        self.add(name + ".weight", (c_out, c_in, k, k), "w")
        if bias:
            self.add(name + ".bias", (c_out,), "b")

    def norm(self, name, n):
        self.add(name + ".weight", (n,), "nw")
        self.add(name + ".bias", (n,), "nb")


# ---------------------------------------------------------------- UNet (unet.c:110-281)
UNET = {
    "sd1":  dict(n_ch=320, ch_mult=[1, 2, 4, 4], attn_res=[4, 2, 1], depth=[1, 1, 1, 1],
                 n_head=8, d_head=0, n_ctx=768, adm=0, linear_proj=False),
    "sd2":  dict(n_ch=320, ch_mult=[1, 2, 4, 4], attn_res=[4, 2, 1], depth=[1, 1, 1, 1],
                 n_head=0, d_head=64, n_ctx=1024, adm=0, linear_proj=True),
    "sdxl": dict(n_ch=320, ch_mult=[1, 2, 4], attn_res=[4, 2], depth=[1, 2, 10],
                 n_head=0, d_head=64, n_ctx=2048, adm=2816, linear_proj=True),
}


def unet_resnet(S, p, c_in, c_out, n_te=1280):
    S.norm(p + "in_layers.0", c_in)
    S.conv(p + "in_layers.2", c_in, c_out, 3)
    S.linear(p + "emb_layers.1", n_te, c_out)
    S.norm(p + "out_layers.0", c_out)
    S.conv(p + "out_layers.3", c_out, c_out, 3)
    if c_in != c_out:
        S.conv(p + "skip_connection", c_in, c_out, 1)


def unet_transf(S, p, ch, depth, n_ctx, linear_proj):
    S.norm(p + "norm", ch)
    if linear_proj:
        S.linear(p + "proj_in", ch, ch)
    else:
        S.conv(p + "proj_in", ch, ch, 1)
    for d in range(depth):
        q = p + "transformer_blocks.%d." % d
        S.norm(q + "norm1", ch)
        for n in ("to_q", "to_k", "to_v"):
            S.linear(q + "attn1." + n, ch, ch, bias=False)
        S.linear(q + "attn1.to_out.0", ch, ch)
        S.norm(q + "norm2", ch)
        S.linear(q + "attn2.to_q", ch, ch, bias=False)
        S.linear(q + "attn2.to_k", n_ctx, ch, bias=False)
        S.linear(q + "attn2.to_v", n_ctx, ch, bias=False)
        S.linear(q + "attn2.to_out.0", ch, ch)
        S.norm(q + "norm3", ch)
        S.linear(q + "ff.net.0.proj", ch, ch * 8)
        S.linear(q + "ff.net.2", ch * 4, ch)
    if linear_proj:
        S.linear(p + "proj_out", ch, ch)
    else:
        S.conv(p + "proj_out", ch, ch, 1)


def spec_unet(S, kind):
    P = UNET[kind]
    n_ch, n_te = P["n_ch"], 1280
    pre = "model.diffusion_model."
    S.linear(pre + "time_embed.0", n_ch, n_te)
    S.linear(pre + "time_embed.2", n_te, n_te)
    if P["adm"]:
        S.linear(pre + "label_emb.0.0", P["adm"], n_te)
        S.linear(pre + "label_emb.0.2", n_te, n_te)
    S.conv(pre + "input_blocks.0.0", 4, n_ch, 3)
    # input blocks (unet.c:167-203)
    skip_ch = [n_ch]
    ch, i_blk, ds = n_ch, 0, 1
    for im, mult in enumerate(P["ch_mult"]):
        if im:
            ds *= 2
            i_blk += 1
            S.conv(pre + "input_blocks.%d.0.op" % i_blk, ch, ch, 3)
            skip_ch.append(ch)
        for _ in range(2):
            i_blk += 1
            c_out = n_ch * mult
            unet_resnet(S, pre + "input_blocks.%d.0." % i_blk, ch, c_out)
            ch = c_out
            if ds in P["attn_res"]:
                unet_transf(S, pre + "input_blocks.%d.1." % i_blk, ch, P["depth"][im], P["n_ctx"], P["linear_proj"])
            skip_ch.append(ch)
    # middle (unet.c:205-217)
    im = len(P["ch_mult"]) - 1
    unet_resnet(S, pre + "middle_block.0.", ch, ch)
    unet_transf(S, pre + "middle_block.1.", ch, P["depth"][im], P["n_ctx"], P["linear_proj"])
    unet_resnet(S, pre + "middle_block.2.", ch, ch)
    # output blocks (unet.c:219-258)
    i_o = 0
    for im in range(len(P["ch_mult"]) - 1, -1, -1):
        for j in range(3):
            c_in = ch + skip_ch.pop()
            c_out = n_ch * P["ch_mult"][im]
            sub = 0
            unet_resnet(S, pre + "output_blocks.%d.%d." % (i_o, sub), c_in, c_out)
            sub += 1
            ch = c_out
            if ds in P["attn_res"]:
                unet_transf(S, pre + "output_blocks.%d.%d." % (i_o, sub), ch, P["depth"][im], P["n_ctx"], P["linear_proj"])
                sub += 1
            if im != 0 and j == 2:
                S.conv(pre + "output_blocks.%d.%d.conv" % (i_o, sub), ch, ch, 3)
                ds //= 2
            i_o += 1
    assert not skip_ch
    S.norm(pre + "out.0", ch)
    S.conv(pre + "out.2", ch, 4, 3)


# ---------------------------------------------------------------- VAE (vae.c:46-180)
def vae_resnet(S, p, c_in, c_out):
    S.norm(p + "norm1", c_in)
    S.conv(p + "conv1", c_in, c_out, 3)
    S.norm(p + "norm2", c_out)
    S.conv(p + "conv2", c_out, c_out, 3)
    if c_in != c_out:
        S.conv(p + "nin_shortcut", c_in, c_out, 1)


def vae_attn(S, p, c):
    S.norm(p + "norm", c)
    for n in ("q", "k", "v", "proj_out"):
        S.conv(p + n, c, c, 1)


def spec_vae(S, encoder=True, decoder=True):
    ch, mult, pre = 128, [1, 2, 4, 4], "first_stage_model."
    if encoder:
        p = pre + "encoder."
        S.conv(p + "conv_in", 3, ch, 3)
        cb = ch
        for i, m in enumerate(mult):
            for j in range(2):
                vae_resnet(S, p + "down.%d.block.%d." % (i, j), cb, ch * m)
                cb = ch * m
            if i + 1 != len(mult):
                S.conv(p + "down.%d.downsample.conv" % i, cb, cb, 3)
        vae_resnet(S, p + "mid.block_1.", cb, cb)
        vae_attn(S, p + "mid.attn_1.", cb)
        vae_resnet(S, p + "mid.block_2.", cb, cb)
        S.norm(p + "norm_out", cb)
        S.conv(p + "conv_out", cb, 8, 3)
        S.conv(pre + "quant_conv", 8, 8, 1)
    if decoder:
        p = pre + "decoder."
        S.conv(pre + "post_quant_conv", 4, 4, 1)
        cb = ch * mult[-1]
        S.conv(p + "conv_in", 4, cb, 3)
        vae_resnet(S, p + "mid.block_1.", cb, cb)
        vae_attn(S, p + "mid.attn_1.", cb)
        vae_resnet(S, p + "mid.block_2.", cb, cb)
        for i in range(len(mult) - 1, -1, -1):
            for j in range(3):
                vae_resnet(S, p + "up.%d.block.%d." % (i, j), cb, ch * mult[i])
                cb = ch * mult[i]
            if i != 0:
                S.conv(p + "up.%d.upsample.conv" % i, cb, cb, 3)
        S.norm(p + "norm_out", cb)
        S.conv(p + "conv_out", cb, 3, 3)


# ---------------------------------------------------------------- CLIP (clip.c:319-437)
CLIP = {
    "l14":  dict(d=768, n_layer=12, n_interm=3072),
    "h14":  dict(d=1024, n_layer=24, n_interm=4096),
    "bigg": dict(d=1280, n_layer=32, n_interm=5120),
}


def spec_clip_hf(S, pre, kind):
    """HF-style names (SD1 cond_stage_model.transformer / SDXL embedders.0)."""
    P = CLIP[kind]
    d = P["d"]
    p = pre + "transformer.text_model."
    S.add(p + "embeddings.token_embedding.weight", (49408, d), "emb")
    S.add(p + "embeddings.position_embedding.weight", (77, d), "pos")
    for l in range(P["n_layer"]):
        q = p + "encoder.layers.%d." % l
        S.norm(q + "layer_norm1", d)
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            S.linear(q + "self_attn." + n, d, d)
        S.norm(q + "layer_norm2", d)
        S.linear(q + "mlp.fc1", d, P["n_interm"])
        S.linear(q + "mlp.fc2", P["n_interm"], d)
    S.norm(p + "final_layer_norm", d)


def spec_clip_open(S, pre, kind):
    """OpenCLIP names with fused in_proj (SD2 cond_stage_model.model / SDXL embedders.1)."""
    P = CLIP[kind]
    d = P["d"]
    p = pre + "model."
    S.add(p + "token_embedding.weight", (49408, d), "emb")
    S.add(p + "positional_embedding", (77, d), "pos")
    for l in range(P["n_layer"]):
        q = p + "transformer.resblocks.%d." % l
        S.norm(q + "ln_1", d)
        S.add(q + "attn.in_proj_weight", (3 * d, d), "w")
        S.add(q + "attn.in_proj_bias", (3 * d,), "b")
        S.linear(q + "attn.out_proj", d, d)
        S.norm(q + "ln_2", d)
        S.linear(q + "mlp.c_fc", d, P["n_interm"])
        S.linear(q + "mlp.c_proj", P["n_interm"], d)
    S.norm(p + "ln_final", d)
    S.add(p + "text_projection", (d, d), "w")


# ---------------------------------------------------------------- TAESD (tae.c:24-92)
def tae_block(S, p, c):
    for i in (0, 2, 4):
        S.conv(p + "conv.%d" % i, c, c, 3)


def spec_tae(S):
    c = 64
    p = "decoder.layers."
    S.conv(p + "0", 4, c, 3)
    i = 2                       # index 1 is the ReLU
    for _ in range(3):
        for _ in range(3):
            tae_block(S, p + "%d." % i, c); i += 1
        i += 1                  # upsample
        S.conv(p + "%d" % i, c, c, 3, bias=False); i += 1
    tae_block(S, p + "%d." % i, c); i += 1
    # the decoder emits the image in [0,1] directly (tae.c:65-92): centre the random-init output at mid-grey so that
    # the parity fixtures are not clamped to black
    S.add(p + "%d.weight" % i, (3, c, 3, 3), "w")
    S.add(p + "%d.bias" % i, (3,), "b_mid")
    p = "encoder.layers."
    S.conv(p + "0", 3, c, 3)
    tae_block(S, p + "1.", c)
    i = 2
    for _ in range(3):
        S.conv(p + "%d" % i, c, c, 3, bias=False); i += 1
        for _ in range(3):
            tae_block(S, p + "%d." % i, c); i += 1
    S.conv(p + "%d" % i, c, 4, 3)


def build_spec(kind, parts=("unet", "vae", "clip")):
    S = Spec()
    if kind == "tae":
        spec_tae(S)
        return S
    if "clip" in parts:
        if kind == "sd1":
            spec_clip_hf(S, "cond_stage_model.", "l14")
        elif kind == "sd2":
            spec_clip_open(S, "cond_stage_model.", "h14")
        else:
            spec_clip_hf(S, "conditioner.embedders.0.", "l14")
            spec_clip_open(S, "conditioner.embedders.1.", "bigg")
    if "vae" in parts:
        spec_vae(S)
    if "unet" in parts:
        spec_unet(S, kind)
    return S


def gen_tensor(name, shape, kind, seed):
    rng = np.random.default_rng([zlib.crc32(name.encode()), seed])
    x = rng.standard_normal(shape, dtype=np.float32)
    if kind == "w":
        fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else shape[0]
        x *= GAIN / np.sqrt(fan_in)
    elif kind == "b":
        x *= 0.02
    elif kind == "b_mid":
        x = 0.5 + 0.02 * x
    elif kind == "nw":
        x = 1.0 + 0.05 * x
    elif kind == "nb":
        x *= 0.05
    elif kind == "emb":
        x *= 0.02
    elif kind == "pos":
        x *= 0.01
    return x


def write_safetensors(path, spec, seed=1234, dtype="f16"):
    npdt = {"f16": np.float16, "f32": np.float32}[dtype]
    stdt = {"f16": "F16", "f32": "F32"}[dtype]
    esz = np.dtype(npdt).itemsize
    header, off = {}, 0
    for name, shape, kind in spec.items:
        n = int(np.prod(shape)) * esz
        header[name] = {"dtype": stdt, "shape": list(shape), "data_offsets": [off, off + n]}
        off += n
    hj = json.dumps(header, separators=(",", ":")).encode()
    hj += b" " * ((8 - len(hj) % 8) % 8)
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(hj)))
        f.write(hj)
        for name, shape, kind in spec.items:
            f.write(gen_tensor(name, shape, kind, seed).astype(npdt).tobytes())
    return off


def write_lora(path, kind, rank=16, alpha=16.0, seed=77, targets=("attn", "ff")):
    """Random LoRA on all attention + feed-forward linears of the UNet (config 4).
    Keys: lora_unet_<ldm path with _>.lora_down.weight / .lora_up.weight / .alpha
    (matched by mlimgsynth.c:1067-1092 + tensor_name_conv.c:6-21)."""
    S = Spec(); spec_unet(S, kind)
    L = Spec()
    for name, shape, k in S.items:
        if not name.endswith(".weight") or len(shape) != 2:
            continue
        if not any(("." + t) in name for t in ("attn1", "attn2", "ff")):
            continue
        base = "lora_unet_" + name[len("model.diffusion_model."):-len(".weight")].replace(".", "_")
        n_out, n_in = shape
        L.add(base + ".lora_down.weight", (rank, n_in), "w")
        L.add(base + ".lora_up.weight", (n_out, rank), "w")
        L.add(base + ".alpha", (), "alpha")
    header, off, blobs = {}, 0, []
    for name, shape, k in L.items:
        if k == "alpha":
            x = np.array(alpha, dtype=np.float16)
        else:
            x = gen_tensor(name, shape, k, seed).astype(np.float16)
        b = x.tobytes()
        header[name] = {"dtype": "F16", "shape": list(shape), "data_offsets": [off, off + len(b)]}
        off += len(b); blobs.append(b)
    hj = json.dumps(header, separators=(",", ":")).encode()
    hj += b" " * ((8 - len(hj) % 8) % 8)
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(hj))); f.write(hj)
        for b in blobs:
            f.write(b)
    return off


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("kind", choices=["sd1", "sd2", "sdxl", "tae"])
    ap.add_argument("out")
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--dtype", default="f16")
    ap.add_argument("--parts", default="unet,vae,clip")
    ap.add_argument("--lora", action="store_true", help="write a LoRA file for this UNet instead")
    a = ap.parse_args()
    if a.lora:
        n = write_lora(a.out, a.kind, seed=a.seed)
    else:
        n = write_safetensors(a.out, build_spec(a.kind, a.parts.split(",")), a.seed, a.dtype)
    print("%s: %.1f MB" % (a.out, n / 1e6))


if __name__ == "__main__":
    main()
