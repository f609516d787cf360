"""Graph-builder helpers for the parity tests: the reference's neural blocks expressed through
the ggml-shaped C ABI, in the same op order as mlblock_nn.c / unet.c / vae.c / clip.c, so that the
planner sees exactly the node sequences the reference emits. Weights are random leaves drawn from
a seeded numpy generator; building the same block twice with the same seed (once per library)
gives identical inputs to the oracle and to the CUDA engine.
"""
import numpy as np
from mlimgsynth_b200.ggml import Graph, GGML_TYPE_F32, SCALE_NEAREST


class B:
    def __init__(self, G: Graph, seed=0, gain=0.7):
        self.G, self.g, self.cc = G, G.g, G.cc
        self.rng = np.random.default_rng(seed)
        self.gain = gain

    # ---- leaves
    def w(self, *shape, fan_in=None, dtype=np.float16):
        fan_in = fan_in or int(np.prod(shape[1:]))
        return self.G.leaf((self.rng.standard_normal(shape) * self.gain / np.sqrt(fan_in)).astype(dtype))

    def bias(self, n):
        return self.G.leaf((self.rng.standard_normal(n) * 0.1).astype(np.float32))

    def inp(self, *shape, scale=1.0, dtype=np.float32):
        return self.G.leaf((self.rng.standard_normal(shape) * scale).astype(dtype))

    def ne(self, t):
        return self.G.shape(t)

    # ---- mlblock_nn.c
    def linear(self, x, n_out, bias=True, wdtype=np.float16):          # :16
        n_in = self.ne(x)[0]
        w = self.w(n_out, n_in, dtype=wdtype)
        x = self.g.ggml_mul_mat(self.cc, w, x)
        if bias:
            x = self.g.ggml_add(self.cc, x, self.bias(n_out))
        return x

    def conv2d(self, x, ch_out, k=3, s=1, p=1, bias=True):             # :31
        ch_in = self.ne(x)[2]
        w = self.w(ch_out, ch_in, k, k)
        x = self.g.ggml_conv_2d(self.cc, w, x, s, s, p, p, 1, 1)
        if bias:
            b = self.g.ggml_reshape_4d(self.cc, self.bias(ch_out), 1, 1, ch_out, 1)
            x = self.g.ggml_add(self.cc, x, b)
        return x

    def layer_norm(self, x, eps=1e-5):                                  # :58
        n = self.ne(x)[0]
        x = self.g.ggml_norm(self.cc, x, eps)
        w = self.G.leaf((1 + 0.1 * self.rng.standard_normal(n)).astype(np.float32))
        x = self.g.ggml_mul(self.cc, x, w)
        return self.g.ggml_add(self.cc, x, self.bias(n))

    def groupnorm32(self, x, eps=1e-6):                                 # :78
        n = self.ne(x)[2]
        x = self.g.ggml_group_norm(self.cc, x, 32, eps)
        w = self.G.leaf((1 + 0.1 * self.rng.standard_normal(n)).astype(np.float32))
        b = self.bias(n)
        w = self.g.ggml_reshape_4d(self.cc, w, 1, 1, n, 1)
        b = self.g.ggml_reshape_4d(self.cc, b, 1, 1, n, 1)
        x = self.g.ggml_mul(self.cc, x, w)
        return self.g.ggml_add(self.cc, x, b)

    def downsample(self, x, ch_out, vae=False):                         # :105
        if vae:
            x = self.g.ggml_pad(self.cc, x, 1, 1, 0, 0)
            return self.conv2d(x, ch_out, 3, 2, 0)
        return self.conv2d(x, ch_out, 3, 2, 1)

    def upsample(self, x, ch_out):                                      # :118
        x = self.g.ggml_upscale(self.cc, x, 2, SCALE_NEAREST)
        return self.conv2d(x, ch_out, 3, 1, 1)

    def resnet(self, x, emb, ch_out):                                   # :129
        x0, ch_in = x, self.ne(x)[2]
        x = self.groupnorm32(x)
        x = self.g.ggml_silu_inplace(self.cc, x)
        x = self.conv2d(x, ch_out)
        if emb is not None:
            e = self.g.ggml_silu(self.cc, emb)
            e = self.linear(e, ch_out)
            ne = self.ne(e)
            e = self.g.ggml_reshape_4d(self.cc, e, 1, 1, ne[0], ne[1])
            x = self.g.ggml_add(self.cc, x, e)
        x = self.groupnorm32(x)
        x = self.g.ggml_silu_inplace(self.cc, x)
        x = self.conv2d(x, ch_out)
        if ch_in != ch_out:
            x0 = self.conv2d(x0, ch_out, 1, 1, 0)
        return self.g.ggml_add(self.cc, x, x0)

    def geglu(self, x, d_out):                                          # :159
        x = self.linear(x, d_out * 2)
        t = self.G.t(x)
        ne, nb = list(t.ne), list(t.nb)
        half = ne[0] // 2
        xv = self.g.ggml_view_4d(self.cc, x, half, ne[1], ne[2], ne[3], nb[1], nb[2], nb[3], 0)
        gv = self.g.ggml_view_4d(self.cc, x, half, ne[1], ne[2], ne[3], nb[1], nb[2], nb[3], half * 4)
        gv = self.g.ggml_cont(self.cc, gv)
        gv = self.g.ggml_gelu_inplace(self.cc, gv)
        return self.g.ggml_mul(self.cc, xv, gv)

    def feed_forward(self, x, d_out, mult=4):                           # :175
        d_in = self.ne(x)[0]
        x = self.geglu(x, d_in * mult)
        return self.linear(x, d_out)

    def attention(self, q, k, v, mask):                                 # ggml_extend.c:200
        d_head = self.ne(q)[0]
        kq = self.g.ggml_mul_mat(self.cc, k, q)
        kq = self.g.ggml_scale_inplace(self.cc, kq, 1.0 / np.sqrt(d_head))
        if mask:
            kq = self.g.ggml_diag_mask_inf_inplace(self.cc, kq, 0)
        kq = self.g.ggml_soft_max_inplace(self.cc, kq)
        return self.g.ggml_mul_mat(self.cc, v, kq)

    def attn_mhead(self, q, k, v, d_out, d_embed, n_head, mask=False, bias=False, bias_out=True):   # :190
        g, cc = self.g, self.cc
        nq1, nq2 = self.ne(q)[1:3]
        nk1, nk2 = self.ne(k)[1:3]
        nv1, nv2 = self.ne(v)[1:3]
        d_head = d_embed // n_head
        q = self.linear(q, d_embed, bias)
        q = g.ggml_reshape_4d(cc, q, d_head, n_head, nq1, nq2)
        q = g.ggml_cont(cc, g.ggml_permute(cc, q, 0, 2, 1, 3))
        q = g.ggml_reshape_3d(cc, q, d_head, nq1, n_head * nq2)
        k = self.linear(k, d_embed, bias)
        k = g.ggml_reshape_4d(cc, k, d_head, n_head, nk1, nk2)
        k = g.ggml_cont(cc, g.ggml_permute(cc, k, 0, 2, 1, 3))
        k = g.ggml_reshape_3d(cc, k, d_head, nk1, n_head * nk2)
        v = self.linear(v, d_embed, bias)
        v = g.ggml_reshape_4d(cc, v, d_head, n_head, nv1, nv2)
        v = g.ggml_cont(cc, g.ggml_permute(cc, v, 1, 2, 0, 3))
        v = g.ggml_reshape_3d(cc, v, nv1, d_head, n_head * nv2)
        v = self.attention(q, k, v, mask)
        v = g.ggml_reshape_4d(cc, v, d_head, nq1, n_head, nq2)
        v = g.ggml_cont(cc, g.ggml_permute(cc, v, 0, 2, 1, 3))
        v = g.ggml_reshape_3d(cc, v, d_embed, nq1, nq2)
        return self.linear(v, d_out, bias_out)

    def basic_transf(self, x, c, d_out, d_embed, n_head):               # :234
        r = x
        x = self.layer_norm(x)
        x = self.attn_mhead(x, x, x, d_out, d_embed, n_head)
        x = self.g.ggml_add(self.cc, x, r)
        r = x
        x = self.layer_norm(x)
        x = self.attn_mhead(x, c, c, d_out, d_embed, n_head)
        x = self.g.ggml_add(self.cc, x, r)
        r = x
        x = self.layer_norm(x)
        x = self.feed_forward(x, d_out)
        return self.g.ggml_add(self.cc, x, r)

    # ---- unet.c:110
    def spatial_transf(self, x, ctx, d_embed, n_head, depth=1):
        g, cc = self.g, self.cc
        x0 = x
        w, h, ch_in, nb = self.ne(x)
        x = self.groupnorm32(x)
        x = self.conv2d(x, d_embed, 1, 1, 0)
        x = g.ggml_cont(cc, g.ggml_permute(cc, x, 1, 2, 0, 3))
        x = g.ggml_reshape_3d(cc, x, d_embed, w * h, nb)
        for _ in range(depth):
            x = self.basic_transf(x, ctx, d_embed, d_embed, n_head)
        x = g.ggml_cont(cc, g.ggml_permute(cc, x, 1, 0, 2, 3))
        x = g.ggml_reshape_4d(cc, x, w, h, d_embed, nb)
        x = self.conv2d(x, ch_in, 1, 1, 0)
        return g.ggml_add(cc, x, x0)

    # ---- vae.c:46
    def attn_2d_self(self, x):
        g, cc = self.g, self.cc
        x0 = x
        x = self.groupnorm32(x)
        w, h, c, n = self.ne(x)
        q = self.conv2d(x, c, 1, 1, 0)
        q = g.ggml_cont(cc, g.ggml_permute(cc, q, 1, 2, 0, 3))
        q = g.ggml_reshape_3d(cc, q, c, h * w, n)
        k = self.conv2d(x, c, 1, 1, 0)
        k = g.ggml_cont(cc, g.ggml_permute(cc, k, 1, 2, 0, 3))
        k = g.ggml_reshape_3d(cc, k, c, h * w, n)
        v = self.conv2d(x, c, 1, 1, 0)
        v = g.ggml_reshape_3d(cc, v, h * w, c, n)
        x = self.attention(q, k, v, False)
        x = g.ggml_cont(cc, g.ggml_permute(cc, x, 1, 0, 2, 3))
        x = g.ggml_reshape_4d(cc, x, w, h, c, n)
        x = self.conv2d(x, c, 1, 1, 0)
        return g.ggml_add(cc, x, x0)

    # ---- clip.c:346-378
    def clip_layer(self, x, d_model, n_head, n_interm, quick=True):
        x0 = x
        x = self.layer_norm(x)
        x = self.attn_mhead(x, x, x, d_model, d_model, n_head, mask=True, bias=True, bias_out=True)
        x0 = x = self.g.ggml_add(self.cc, x0, x)
        x = self.layer_norm(x)
        x = self.linear(x, n_interm)
        x = (self.g.ggml_gelu_quick_inplace if quick else self.g.ggml_gelu_inplace)(self.cc, x)
        x = self.linear(x, d_model)
        return self.g.ggml_add(self.cc, x0, x)


def run_both(build_fn, ref, eng, seed=0):
    """build_fn(B) -> output tensor(s). Runs the same graph on the oracle and on the engine."""
    outs = []
    for lib in (ref, eng):
        G = Graph(lib)
        o = build_fn(B(G, seed))
        if not isinstance(o, (list, tuple)):
            o = [o]
        res = G.run(*o)
        G.free()
        outs.append([r.astype(np.float32) for r in res])
    return outs


def max_rel_err(a, b):
    """max |a-b| / max |b|  (the per-step UNet tolerance of BASELINE.json north_star: 1e-2)."""
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-12))
