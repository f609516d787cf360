"""Whole-model anchor for the UNet: one complete denoising step of the reference's own path
(mlimgsynth_cpu generate: tokenizer -> clip.c -> Philox noise -> sampling.c / solvers.c Euler step -> unet.c graph) on the
CPU oracle, against a PyTorch fp32 evaluation of the CompVis latent-diffusion UNet
(ldm.modules.diffusionmodules.openaimodel.UNetModel with SpatialTransformer blocks, the architecture of the
`model.diffusion_model.*` keys) fed by Hugging Face's CLIPTextModel, on the same random-init SD1.x checkpoint.

With one Euler step from sigma_max to 0 and cfg 1 the final latent is the denoised estimate
    x0 - sigma0 * eps(x0 / sqrt(sigma0^2 + 1), t = 999, cond),   x0 = sigma0 * philox_randn(seed)
so the comparison covers the time embedding, every resnet / transformer / up- and down-sampling block, the skip
concatenations, the sigma <-> t mapping, the input scaling and the sampler arithmetic. The torch side is written with
torch.nn.functional only and shares no code with oracle/ggml_ref.c or the reference.

CPU only; uses the 1.7 GB SD1 checkpoint of tools/gen_weights.py (cached under /tmp by bench.weights_path)."""
import ctypes as C
import json, os, struct, subprocess, sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
PROMPT = "a photograph of an astronaut riding a horse"
TOKENS = [320, 8853, 539, 550, 18376, 6765, 320, 4558]          # checked against the reference tokenizer in tests/test_clip_vs_hf.py
SIGMA_MAX = 14.614641                                           # unet.c:36-37


def read_safetensors(path, prefix):
    import torch
    out = {}
    with open(path, "rb") as f:
        n = struct.unpack("<Q", f.read(8))[0]
        hdr = json.loads(f.read(n)); base = 8 + n
        for k, v in hdr.items():
            if k.startswith(prefix):
                f.seek(base + v["data_offsets"][0])
                raw = f.read(v["data_offsets"][1] - v["data_offsets"][0])
                out[k[len(prefix):]] = torch.from_numpy(np.frombuffer(raw, dtype=np.float16).reshape(v["shape"]).astype(np.float32))
    return out


class LdmUnet:
    """model_channels 320, channel_mult (1, 2, 4, 4), 2 res blocks, attention at ds 1 / 2 / 4.
    SD1.x: 8 heads, 1x1-conv proj_in / proj_out, context 768. SD2.x: 64-wide heads, linear proj_in / proj_out, context 1024."""
    def __init__(self, W, n_head=8, d_head=0, linear_proj=False):
        import torch, torch.nn.functional as F
        self.W, self.t, self.F, self.n_head, self.d_head, self.linear_proj = W, torch, F, n_head, d_head, linear_proj

    def lin(self, x, name, bias=True):
        y = x.half().float() @ self.W[name + ".weight"].t()
        return y + self.W[name + ".bias"] if bias else y

    def conv(self, x, name, stride=1):
        w = self.W[name + ".weight"]
        return self.F.conv2d(x.half().float(), w, self.W[name + ".bias"], stride=stride, padding=1 if w.shape[-1] == 3 else 0)

    def gn(self, x, name):
        return self.F.group_norm(x, 32, self.W[name + ".weight"], self.W[name + ".bias"], eps=1e-6)

    def ln(self, x, name):
        return self.F.layer_norm(x, (x.shape[-1],), self.W[name + ".weight"], self.W[name + ".bias"], eps=1e-5)

    def res(self, x, emb, p):
        h = self.conv(self.F.silu(self.gn(x, p + "in_layers.0")), p + "in_layers.2")
        h = h + self.lin(self.F.silu(emb), p + "emb_layers.1")[:, :, None, None]
        h = self.conv(self.F.silu(self.gn(h, p + "out_layers.0")), p + "out_layers.3")
        if (p + "skip_connection.weight") in self.W:
            x = self.conv(x, p + "skip_connection")
        return x + h

    def attention(self, x, ctx, p):
        H = self.n_head if not self.d_head else x.shape[-1] // self.d_head
        q, k, v = self.lin(x, p + "to_q", False), self.lin(ctx, p + "to_k", False), self.lin(ctx, p + "to_v", False)
        b, n, c = q.shape
        split = lambda t: t.reshape(b, t.shape[1], H, c // H).transpose(1, 2)
        w = self.t.softmax(split(q) @ split(k).transpose(-1, -2) * (c // H) ** -0.5, dim=-1)
        o = (w @ split(v)).transpose(1, 2).reshape(b, n, c)
        return self.lin(o, p + "to_out.0")

    def transformer(self, x, ctx, p):
        b, c, hh, ww = x.shape
        if self.linear_proj:
            h = self.lin(self.gn(x, p + "norm").reshape(b, c, hh * ww).transpose(1, 2), p + "proj_in")
        else:
            h = self.conv(self.gn(x, p + "norm"), p + "proj_in").reshape(b, c, hh * ww).transpose(1, 2)
        q = p + "transformer_blocks.0."
        h = h + self.attention(self.ln(h, q + "norm1"), self.ln(h, q + "norm1"), q + "attn1.")
        h = h + self.attention(self.ln(h, q + "norm2"), ctx, q + "attn2.")
        g = self.lin(self.ln(h, q + "norm3"), q + "ff.net.0.proj")
        val, gate = g.chunk(2, dim=-1)
        h = h + self.lin(val * self.F.gelu(gate, approximate="tanh"), q + "ff.net.2")
        if self.linear_proj:
            return x + self.lin(h, p + "proj_out").transpose(1, 2).reshape(b, c, hh, ww)
        return x + self.conv(h.transpose(1, 2).reshape(b, c, hh, ww), p + "proj_out")

    def forward(self, x, t, ctx):
        torch, F = self.t, self.F
        half = 160
        f = torch.exp(-np.log(10000.0) * torch.arange(half, dtype=torch.float32) / half)
        emb = torch.cat([torch.cos(t * f), torch.sin(t * f)])[None]
        emb = self.lin(F.silu(self.lin(emb, "time_embed.0")), "time_embed.2")
        hs = []
        h = self.conv(x, "input_blocks.0.0"); hs.append(h)
        i_blk, ds = 0, 1
        for im in range(4):
            if im:
                ds *= 2; i_blk += 1
                h = self.conv(h, "input_blocks.%d.0.op" % i_blk, stride=2); hs.append(h)
            for _ in range(2):
                i_blk += 1
                h = self.res(h, emb, "input_blocks.%d.0." % i_blk)
                if ds in (1, 2, 4):
                    h = self.transformer(h, ctx, "input_blocks.%d.1." % i_blk)
                hs.append(h)
        h = self.res(h, emb, "middle_block.0.")
        h = self.transformer(h, ctx, "middle_block.1.")
        h = self.res(h, emb, "middle_block.2.")
        i_o = 0
        for im in (3, 2, 1, 0):
            for j in range(3):
                h = self.res(torch.cat([h, hs.pop()], 1), emb, "output_blocks.%d.0." % i_o)
                sub = 1
                if ds in (1, 2, 4):
                    h = self.transformer(h, ctx, "output_blocks.%d.1." % i_o); sub = 2
                if im and j == 2:
                    h = self.conv(F.interpolate(h, scale_factor=2, mode="nearest"), "output_blocks.%d.%d.conv" % (i_o, sub))
                    ds //= 2
                i_o += 1
        assert not hs
        return self.conv(F.silu(self.gn(h, "out.0")), "out.2")


def load_tensor(path):
    with open(path, "rb") as f:
        head = f.readline().split()
        ne = [int(x) for x in head[2:6]]
        return np.frombuffer(f.read(), dtype=np.float32).reshape(ne[::-1])


@pytest.mark.parametrize("kind", ["sd1", "sd2"])
def test_reference_denoising_step_on_oracle_matches_ldm_unet(oracle_built, tmp_path, kind):
    """sd1: eps-prediction, CLIP ViT-L/14 (last block). sd2: v-prediction (unet.c:490-494), OpenCLIP ViT-H/14 (penultimate
    block + ln_final), 64-wide heads, linear projections."""
    torch = pytest.importorskip("torch")
    tr = pytest.importorskip("transformers")
    import bench, mlimgsynth_b200
    from test_clip_vs_hf import openclip_to_hf
    exe = os.path.join(oracle_built, "mlimgsynth_cpu")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/mlimgsynth_cpu not built")
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    wpath = bench.weights_path(kind)
    seed = 42

    # ---- the reference's own path on the oracle: one Euler step, no guidance, 128x128
    out = str(tmp_path / "o")
    r = subprocess.run([exe, "generate", "-m", wpath, "-p", PROMPT, "-d", "128,128", "-S", str(seed), "-s", "1", "--method", "euler",
                        "--cfg-scale", "1", "-o", out + ".pnm", "--olatent", out + ".tensor"], capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stderr[-1500:]
    got = load_tensor(out + ".tensor")[0]                                     # [4, 16, 16]

    # ---- PyTorch: HF CLIP text encoder -> LDM UNet -> the same Euler step
    if kind == "sd1":
        cfg = tr.CLIPTextConfig(vocab_size=49408, hidden_size=768, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=12,
                                max_position_embeddings=77, hidden_act="quick_gelu")
        sd, pad = read_safetensors(wpath, "cond_stage_model.transformer."), 49407
    else:
        cfg = tr.CLIPTextConfig(vocab_size=49408, hidden_size=1024, intermediate_size=4096, num_hidden_layers=24, num_attention_heads=16,
                                max_position_embeddings=77, hidden_act="gelu_pytorch_tanh")
        sd, pad = openclip_to_hf(read_safetensors(wpath, "cond_stage_model.model."), 1024), 0
    clip = tr.CLIPTextModel(cfg).eval()
    missing, unexpected = clip.load_state_dict(sd, strict=False)
    assert not missing and not unexpected
    ids = [49406] + TOKENS + [49407]
    ids += [pad] * (77 - len(ids))

    class Rng(C.Structure):
        _fields_ = [("seed", C.c_uint64), ("offset", C.c_uint32)]
    L = C.CDLL(mlimgsynth_b200.HOST_LIB)
    buf = (C.c_float * 1024)()
    L.rng_philox_randn(C.byref(Rng(seed, 0)), 1024, buf)                       # bit-exact with the reference's stream (test_host_cpu.py)
    noise = torch.from_numpy(np.frombuffer(buf, dtype=np.float32).reshape(1, 4, 16, 16).copy())
    W = read_safetensors(wpath, "model.diffusion_model.")
    unet = LdmUnet(W) if kind == "sd1" else LdmUnet(W, n_head=0, d_head=64, linear_proj=True)
    with torch.no_grad():
        o = clip(input_ids=torch.tensor([ids]), output_hidden_states=True)
        cond = o.last_hidden_state if kind == "sd1" else clip.text_model.final_layer_norm(o.hidden_states[-2])
        x0 = noise * SIGMA_MAX
        net = unet.forward(x0 / np.sqrt(SIGMA_MAX ** 2 + 1), 999.0, cond)
        if kind == "sd1":
            dx = net                                                           # eps-prediction: dx/dsigma = eps
        else:
            dx = net / np.sqrt(SIGMA_MAX ** 2 + 1) + x0 * (SIGMA_MAX / (SIGMA_MAX ** 2 + 1))      # v-prediction (unet.c:490-494)
        want = (x0 - SIGMA_MAX * dx)[0].numpy()
    err = np.abs(got - want).max() / np.abs(want).max()
    # the latent is dominated by x0; the derivative the sampler used, recovered from the step: dx = (x0 - latent) / sigma0
    dx_got, dx_want = (x0[0].numpy() - got) / SIGMA_MAX, dx[0].numpy()
    err_dx = np.abs(dx_got - dx_want).max() / np.abs(dx_want).max()
    print("%s: one Euler step of the reference on the oracle vs HF CLIP + LDM UNet in torch: latent max-rel err %.2e, dx (max |dx| %.2f) "
          "max-rel err %.2e" % (kind, err, np.abs(dx_want).max(), err_dx))
    # (sd2: the step's result is x0 / (sigma0^2 + 1) - v * sigma0 / sqrt(sigma0^2 + 1), i.e. essentially -v: there the latent error IS the UNet output error)
    assert err <= 3e-3 and err_dx <= 1e-2          # the north-star tolerance for a UNet output is 1e-2
