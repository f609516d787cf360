"""A second, independent anchor for the CPU oracle (oracle/ggml_ref.c, the restatement of ggml's operator semantics).

ggml itself is not on this machine and the reference's tests pin no numeric graph output, so the
arithmetic of the oracle is checked here against PyTorch's fp32 CPU operators: the Stable Diffusion checkpoints were
trained with exactly these operators, which is what any correct ggml build has to reproduce. Every test builds the
reference's op sequence through the ggml-shaped C ABI (tests/blocks.py, same order as mlblock_nn.c / unet.c), runs it
on the oracle, and evaluates the same function with torch on the same leaves. Activations entering an f16-weight
mul_mat / conv are rounded to f16 first, as ggml's CPU path does; tolerances state what is left.

CPU only (no GPU, no engine): runs in the `-m "not gpu"` suite.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from mlimgsynth_b200.ggml import Graph
from blocks import B

torch.set_num_threads(4)
TOL = 2e-5          # pure f32 ops: summation order only
TOL_H = 1e-3        # ops whose operands ggml rounds to f16


def run_oracle(ref, build, seed=0):
    """-> (outputs as float32 numpy arrays, leaves in creation order as numpy arrays)"""
    G = Graph(ref)
    b = B(G, seed)
    o = build(b)
    res = G.run(o)[0].astype(np.float32)
    leaves = [arr.copy() for _, arr in G.leaves]
    G.free()
    return res, leaves


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a).astype(np.float32))


def r16(t):
    return t.half().float()


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-12))


def test_linear_matches_torch(ref):
    out, (x, w, bias) = run_oracle(ref, lambda b: b.linear(b.inp(37, 320), 192))
    want = r16(T(x)) @ T(w).t() + T(bias)
    assert rel(out, want.numpy()) <= TOL_H


@pytest.mark.parametrize("k,s,p", [(3, 1, 1), (3, 2, 1), (1, 1, 0)])
def test_conv2d_matches_torch(ref, k, s, p):
    out, (x, w, bias) = run_oracle(ref, lambda b: b.conv2d(b.inp(2, 24, 13, 10), 40, k, s, p))
    want = F.conv2d(r16(T(x)), T(w), T(bias), stride=s, padding=p)
    assert out.shape == tuple(want.shape)
    assert rel(out, want.numpy()) <= TOL_H


def test_vae_downsample_pads_right_and_bottom(ref):
    """mlb_downsample with the VAE flag: ggml_pad adds one column / row at the END of each axis, then a stride-2 conv
    without padding (mlblock_nn.c:105-116)."""
    out, (x, w, bias) = run_oracle(ref, lambda b: b.downsample(b.inp(1, 16, 10, 12), 24, vae=True))
    want = F.conv2d(F.pad(r16(T(x)), (0, 1, 0, 1)), T(w), T(bias), stride=2)
    assert out.shape == tuple(want.shape) and rel(out, want.numpy()) <= TOL_H


def test_groupnorm_matches_torch(ref):
    """32 groups, eps 1e-6 (mlblock_nn.h:24), affine per channel."""
    out, (x, w, bias) = run_oracle(ref, lambda b: b.groupnorm32(b.inp(2, 64, 9, 7, scale=2.0)))
    want = F.group_norm(T(x), 32, T(w), T(bias), eps=1e-6)
    assert rel(out, want.numpy()) <= TOL


def test_layernorm_matches_torch(ref):
    out, (x, w, bias) = run_oracle(ref, lambda b: b.layer_norm(b.inp(19, 320, scale=3.0)))
    want = F.layer_norm(T(x), (320,), T(w), T(bias), eps=1e-5)
    assert rel(out, want.numpy()) <= TOL


@pytest.mark.parametrize("op", ["silu", "gelu", "gelu_quick", "relu", "tanh"])
def test_unary_matches_torch(ref, op):
    def build(b):
        x = b.g.ggml_scale(b.cc, b.inp(7, 33, scale=3.0), 1.5)
        f = {"silu": b.g.ggml_silu_inplace, "gelu": b.g.ggml_gelu_inplace, "gelu_quick": b.g.ggml_gelu_quick_inplace,
             "relu": b.g.ggml_relu_inplace, "tanh": b.g.ggml_tanh_inplace}[op]
        return f(b.cc, x)
    out, (x,) = run_oracle(ref, build)
    t = T(x) * 1.5
    want = {"silu": F.silu, "gelu": lambda v: F.gelu(v, approximate="tanh"), "gelu_quick": lambda v: v * torch.sigmoid(1.702 * v),
            "relu": F.relu, "tanh": torch.tanh}[op](t)
    # ggml evaluates gelu / gelu_quick / silu through f16 lookup tables or f16-rounded inputs on some builds: allow f16 resolution
    assert rel(out, want.numpy()) <= TOL_H


def test_softmax_matches_torch(ref):
    out, (x,) = run_oracle(ref, lambda b: b.g.ggml_soft_max_inplace(b.cc, b.g.ggml_scale(b.cc, b.inp(3, 9, 50, scale=2.0), 1.0)))
    assert rel(out, torch.softmax(T(x), -1).numpy()) <= TOL


def test_timestep_embedding_is_cos_first(ref):
    """ggml_timestep_embedding: [cos(t f_i) | sin(t f_i)], f_i = exp(-ln(max_period) i / half) -- the CompVis order the
    reference relies on (unet.c:147-170)."""
    t0 = 981.3
    out, _ = run_oracle(ref, lambda b: b.g.ggml_timestep_embedding(b.cc, b.G.leaf(np.array([t0], dtype=np.float32)), 320, 10000))
    half = 160
    f = torch.exp(-np.log(10000.0) * torch.arange(half, dtype=torch.float32) / half)
    want = torch.cat([torch.cos(t0 * f), torch.sin(t0 * f)])
    assert rel(out.reshape(-1), want.numpy()) <= 1e-4


def test_concat_pad_upscale_match_torch(ref):
    def build(b):
        x = b.inp(2, 8, 6, 5); y = b.inp(2, 16, 6, 5)
        z = b.g.ggml_concat(b.cc, x, y, 2)                 # channels
        z = b.g.ggml_pad(b.cc, z, 1, 1, 0, 0)              # one column / row at the end
        return b.g.ggml_upscale(b.cc, z, 2, 0)             # nearest
    out, (x, y) = run_oracle(ref, build)
    want = F.interpolate(F.pad(torch.cat([T(x), T(y)], 1), (0, 1, 0, 1)), scale_factor=2, mode="nearest")
    assert out.shape == tuple(want.shape) and rel(out, want.numpy()) <= 1e-6


@pytest.mark.parametrize("mask", [False, True])
def test_multihead_attention_matches_torch(ref, mask):
    """mlb_attn_mhead (mlblock_nn.c:190-231): projections, head split by permute, softmax(q k^T / sqrt(d)) v with an optional
    causal mask (ggml_diag_mask_inf), head merge, output projection."""
    nq, d, H = 21, 64, 4
    def build(b):
        x = b.inp(nq, d)
        return b.attn_mhead(x, x, x, d, d, H, mask=mask, bias=True)
    out, lv = run_oracle(ref, build)
    x, wq, bq, wk, bk, wv, bv, wo, bo = [T(a) for a in lv]
    lin = lambda t, w, bb: r16(t) @ w.t() + bb
    q = lin(x, wq, bq).view(nq, H, d // H).transpose(0, 1)
    k = lin(x, wk, bk).view(nq, H, d // H).transpose(0, 1)
    v = lin(x, wv, bv).view(nq, H, d // H).transpose(0, 1)
    s = (q @ k.transpose(1, 2)) / np.sqrt(d // H)          # activations x activations: f32 in the reference graph
    if mask:
        s = s + torch.triu(torch.full((nq, nq), float("-inf")), 1)
    o = (torch.softmax(s, -1) @ v).transpose(0, 1).reshape(nq, d)
    want = lin(o, wo, bo)
    assert rel(out, want.numpy()) <= 2e-3


def test_resnet_block_matches_torch(ref):
    """mlb_resnet (mlblock_nn.c:129-157): GN+SiLU, conv, + linear(silu(emb)) per image, GN+SiLU, conv, skip (1x1 when widths differ)."""
    def build(b):
        return b.resnet(b.inp(2, 32, 8, 8), b.inp(2, 96), 64)
    out, lv = run_oracle(ref, build)
    x, emb, g1w, g1b, c1w, c1b, ew, eb, g2w, g2b, c2w, c2b, sw, sb = [T(a) for a in lv]
    h = F.silu(F.group_norm(x, 32, g1w, g1b, eps=1e-6))
    h = F.conv2d(r16(h), c1w, c1b, padding=1)
    e = r16(F.silu(emb)) @ ew.t() + eb
    h = h + e[:, :, None, None]
    h = F.silu(F.group_norm(h, 32, g2w, g2b, eps=1e-6))
    h = F.conv2d(r16(h), c2w, c2b, padding=1)
    want = h + F.conv2d(r16(x), sw, sb)
    assert rel(out, want.numpy()) <= 2e-3


def test_geglu_feed_forward_matches_torch(ref):
    """mlb_GEGLU + linear (mlblock_nn.c:159-188): proj -> (value | gate) halves -> value * gelu_tanh(gate) -> linear."""
    out, lv = run_oracle(ref, lambda b: b.feed_forward(b.inp(11, 64), 64))
    x, w1, b1, w2, b2 = [T(a) for a in lv]
    h = r16(x) @ w1.t() + b1
    val, gate = h[:, :256], h[:, 256:]
    want = r16(val * F.gelu(gate, approximate="tanh")) @ w2.t() + b2
    assert rel(out, want.numpy()) <= 2e-3
