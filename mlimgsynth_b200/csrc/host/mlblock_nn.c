/* mlblock_nn.c -- see mlblock_nn.h. Each block records the op sequence the engine's planner fuses:
 *   linear/conv + bias (+ emb add, activation, residual)  -> one tcgen05 GEMM with epilogue
 *   group_norm * w + b (+ silu), norm * w + b              -> one normalisation kernel
 *   mul_mat, scale, [mask], soft_max, mul_mat              -> one attention kernel
 *   chunk views, cont, gelu, mul                           -> one GEGLU gate kernel
 * Parameter names follow the reference (mlblock_nn.c:16-253) so LDM checkpoints load unchanged. */
#include "mlblock_nn.h"
#include <math.h>

#define N(name, x) mlctx_tensor_add(C, (name), (x))

MLTensor* mlb_nn_linear(MLCtx* C, MLTensor* x, int n_out, bool bias)
{
	mlctx_block_begin(C);
	MLTensor* w = N("weight", ggml_new_tensor_2d(C->cp, C->c.wtype, x->ne[0], n_out));
	x = ggml_mul_mat(C->cc, w, x);
	if (bias) x = ggml_add(C->cc, x, N("bias", ggml_new_tensor_1d(C->cp, GGML_TYPE_F32, n_out)));
	return x;
}

MLTensor* mlb_nn_conv2d(MLCtx* C, MLTensor* x, int ch_out, int k0, int k1, int s0, int s1, int p0, int p1, int d0, int d1, bool bias)
{
	mlctx_block_begin(C);
	/* convolution kernels are always half precision (mlblock_nn.c:43) */
	MLTensor* w = N("weight", ggml_new_tensor_4d(C->cp, GGML_TYPE_F16, k0, k1, x->ne[2], ch_out));
	x = ggml_conv_2d(C->cc, w, x, s0, s1, p0, p1, d0, d1);
	if (bias) {
		MLTensor* b = N("bias", ggml_new_tensor_1d(C->cp, GGML_TYPE_F32, ch_out));
		x = ggml_add(C->cc, x, ggml_reshape_4d(C->cc, b, 1, 1, ch_out, 1));
	}
	return x;
}

MLTensor* mlb_nn_layer_norm(MLCtx* C, MLTensor* x, bool affine, bool bias, float eps)
{
	mlctx_block_begin(C);
	int n = (int)x->ne[0];
	x = ggml_norm(C->cc, x, eps > 0 ? eps : 1e-5f);
	if (affine) {
		x = ggml_mul(C->cc, x, N("weight", ggml_new_tensor_1d(C->cp, GGML_TYPE_F32, n)));
		if (bias) x = ggml_add(C->cc, x, N("bias", ggml_new_tensor_1d(C->cp, GGML_TYPE_F32, n)));
	}
	return x;
}

MLTensor* mlb_nn_groupnorm(MLCtx* C, MLTensor* x, int n_grp, bool affine, float eps)
{
	mlctx_block_begin(C);
	int n = (int)x->ne[2];
	x = ggml_group_norm(C->cc, x, n_grp, eps > 0 ? eps : 1e-5f);
	if (affine) {
		MLTensor* w = N("weight", ggml_new_tensor_1d(C->cp, GGML_TYPE_F32, n));
		MLTensor* b = N("bias", ggml_new_tensor_1d(C->cp, GGML_TYPE_F32, n));
		x = ggml_mul(C->cc, x, ggml_reshape_4d(C->cc, w, 1, 1, n, 1));
		x = ggml_add(C->cc, x, ggml_reshape_4d(C->cc, b, 1, 1, n, 1));
	}
	return x;
}

MLTensor* mlb_downsample(MLCtx* C, MLTensor* x, int ch_out, bool vae)
{
	mlctx_block_begin(C);
	if (vae) {   /* asymmetric: zero-pad right/bottom, then stride 2 without padding */
		x = ggml_pad(C->cc, x, 1, 1, 0, 0);
		return N("conv", mlb_nn_conv2d(C, x, ch_out, 3, 3, 2, 2, 0, 0, 1, 1, true));
	}
	return N("conv", mlb_nn_conv2d(C, x, ch_out, 3, 3, 2, 2, 1, 1, 1, 1, true));
}

MLTensor* mlb_upsample(MLCtx* C, MLTensor* x, int ch_out)
{
	mlctx_block_begin(C);
	x = ggml_upscale(C->cc, x, 2, GGML_SCALE_MODE_NEAREST);
	return N("conv", mlb_nn_conv2d(C, x, ch_out, 3, 3, 1, 1, 1, 1, 1, 1, true));
}

MLTensor* mlb_resnet(MLCtx* C, MLTensor* x, MLTensor* emb, int ch_out)
{
	MLTensor* skip = x;
	int ch_in = (int)x->ne[2];
	mlctx_block_begin(C);
	x = ggml_silu_inplace(C->cc, N("norm1", mlb_nn_groupnorm32(C, x)));
	x = N("conv1", mlb_nn_conv2d(C, x, ch_out, 3, 3, 1, 1, 1, 1, 1, 1, true));
	if (emb) {   /* per-image time/label embedding, broadcast over the pixels */
		MLTensor* e = N("emb_proj", mlb_nn_linear(C, ggml_silu(C->cc, emb), ch_out, true));
		x = ggml_add(C->cc, x, ggml_reshape_4d(C->cc, e, 1, 1, e->ne[0], e->ne[1]));
	}
	x = ggml_silu_inplace(C->cc, N("norm2", mlb_nn_groupnorm32(C, x)));
	x = N("conv2", mlb_nn_conv2d(C, x, ch_out, 3, 3, 1, 1, 1, 1, 1, 1, true));
	if (ch_in != ch_out) skip = N("skip_conv", mlb_nn_conv2d(C, skip, ch_out, 1, 1, 1, 1, 0, 0, 1, 1, true));
	return ggml_add(C->cc, x, skip);
}

MLTensor* mlb_GEGLU(MLCtx* C, MLTensor* x, int d_out)
{
	mlctx_block_begin(C);
	x = N("proj", mlb_nn_linear(C, x, d_out * 2, true));
	/* value = first half of dim 0, gate = second half (strided views, ggml_extend.c:135-155) */
	size_t es = ggml_element_size(x);
	MLTensor* val  = ggml_view_4d(C->cc, x, d_out, x->ne[1], x->ne[2], x->ne[3], x->nb[1], x->nb[2], x->nb[3], 0);
	MLTensor* gate = ggml_view_4d(C->cc, x, d_out, x->ne[1], x->ne[2], x->ne[3], x->nb[1], x->nb[2], x->nb[3], es * d_out);
	gate = ggml_gelu_inplace(C->cc, ggml_cont(C->cc, gate));
	return ggml_mul(C->cc, val, gate);
}

MLTensor* mlb_feed_forward(MLCtx* C, MLTensor* x, int d_out, int mult)
{
	mlctx_block_begin(C);
	int d_inner = (int)x->ne[0] * mult;
	x = N("net.0", mlb_GEGLU(C, x, d_inner));
	return N("net.2", mlb_nn_linear(C, x, d_out, true));
}

MLTensor* mlb_attention(MLCtx* C, MLTensor* q, MLTensor* k, MLTensor* vt, bool mask)
{
	MLTensor* kq = ggml_mul_mat(C->cc, k, q);
	kq = ggml_scale_inplace(C->cc, kq, 1.0f / sqrtf((float)q->ne[0]));
	if (mask) kq = ggml_diag_mask_inf_inplace(C->cc, kq, 0);
	kq = ggml_soft_max_inplace(C->cc, kq);
	return ggml_mul_mat(C->cc, vt, kq);
}

/* x: [d_embed, n_tok, n_batch] -> heads split off dim 0: [d_head, n_tok, n_head, n_batch] */
static MLTensor* split_heads(MLCtx* C, MLTensor* x, int d_head, int n_head)
{
	x = ggml_reshape_4d(C->cc, x, d_head, n_head, x->ne[1], x->ne[2]);
	return ggml_cont(C->cc, ggml_permute(C->cc, x, 0, 2, 1, 3));
}

MLTensor* mlb_attn_mhead(MLCtx* C, MLTensor* q, MLTensor* k, MLTensor* v, int d_out, int d_embed, int n_head, bool mask, bool bias, bool bias_out)
{
	GGML_ASSERT(q->ne[3] == 1 && k->ne[3] == 1 && v->ne[3] == 1);
	int d_head = d_embed / n_head;
	GGML_ASSERT(d_head * n_head == d_embed);
	int64_t nq = q->ne[1], nb = q->ne[2];
	mlctx_block_begin(C);
	q = split_heads(C, N("q_proj", mlb_nn_linear(C, q, d_embed, bias)), d_head, n_head);
	k = split_heads(C, N("k_proj", mlb_nn_linear(C, k, d_embed, bias)), d_head, n_head);
	v = N("v_proj", mlb_nn_linear(C, v, d_embed, bias));
	v = ggml_reshape_4d(C->cc, v, d_head, n_head, v->ne[1], v->ne[2]);
	v = ggml_cont(C->cc, ggml_permute(C->cc, v, 1, 2, 0, 3));          /* [n_tok, d_head, n_head, n_batch] */
	MLTensor* o = mlb_attention(C, q, k, v, mask);                      /* [d_head, nq, n_head, n_batch] */
	o = ggml_cont(C->cc, ggml_permute(C->cc, o, 0, 2, 1, 3));          /* [d_head, n_head, nq, n_batch] */
	o = ggml_reshape_3d(C->cc, o, d_embed, nq, nb);
	return N("out_proj", mlb_nn_linear(C, o, d_out, bias_out));
}

MLTensor* mlb_basic_transf(MLCtx* C, MLTensor* x, MLTensor* c, int d_out, int d_embed, int n_head)
{
	mlctx_block_begin(C);
	MLTensor* h = N("norm1", mlb_nn_layer_norm(C, x, true, true, 0));
	x = ggml_add(C->cc, N("attn1", mlb_attn_mhead(C, h, h, h, d_out, d_embed, n_head, false, false, true)), x);
	h = N("norm2", mlb_nn_layer_norm(C, x, true, true, 0));
	x = ggml_add(C->cc, N("attn2", mlb_attn_mhead(C, h, c, c, d_out, d_embed, n_head, false, false, true)), x);
	h = N("norm3", mlb_nn_layer_norm(C, x, true, true, 0));
	return ggml_add(C->cc, N("ff", mlb_feed_forward(C, h, d_out, 4)), x);
}
