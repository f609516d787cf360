#include "unet.h"
#include "mlblock_nn.h"
#include <math.h>

#define N(name, x) mlctx_tensor_add(C, (name), (x))

static float g_log_sigmas_sd[1000];

/* Architecture tables (unet.c:21-80) */
const UnetParams g_unet_sd1 = {
	.n_ch_in = 4, .n_ch_out = 4, .n_res_blk = 2, .attn_res = {4, 2, 1}, .ch_mult = {1, 2, 4, 4}, .transf_depth = {1, 1, 1, 1},
	.n_te = 1280, .n_head = 8, .n_ctx = 768, .n_ch = 320, .clip_norm = 1,
	.n_step_train = 1000, .sigma_min = 0.029167158f, .sigma_max = 14.614641f, .log_sigmas = g_log_sigmas_sd };
const UnetParams g_unet_sd2 = {
	.n_ch_in = 4, .n_ch_out = 4, .n_res_blk = 2, .attn_res = {4, 2, 1}, .ch_mult = {1, 2, 4, 4}, .transf_depth = {1, 1, 1, 1},
	.n_te = 1280, .d_head = 64, .n_ctx = 1024, .n_ch = 320, .clip_norm = 1, .vparam = 1,
	.n_step_train = 1000, .sigma_min = 0.029167158f, .sigma_max = 14.614641f, .log_sigmas = g_log_sigmas_sd };
const UnetParams g_unet_sdxl = {
	.n_ch_in = 4, .n_ch_out = 4, .n_res_blk = 2, .attn_res = {4, 2}, .ch_mult = {1, 2, 4}, .transf_depth = {1, 2, 10},
	.n_te = 1280, .d_head = 64, .n_ctx = 2048, .n_ch = 320, .ch_adm_in = 2816, .cond_label = 1, .uncond_empty_zero = 1,
	.n_step_train = 1000, .sigma_min = 0.029167158f, .sigma_max = 14.614641f, .log_sigmas = g_log_sigmas_sd };

static bool in_list(const int* v, int x) { for (; *v; ++v) if (*v == x) return true; return false; }

/* Spatial transformer (unet.c:110-145): GN, 1x1 in, tokens, depth x basic_transf, back, 1x1 out, residual.
 * The two permute+cont pairs are free views in the engine (channels-last activations). */
static MLTensor* mlb_spatial_transf(MLCtx* C, MLTensor* x, MLTensor* ctx, int d_embed, int d_head, int n_head, int depth)
{
	MLTensor* skip = x;
	char name[32];
	mlctx_block_begin(C);
	int64_t w = x->ne[0], h = x->ne[1], ch_in = x->ne[2], nb = x->ne[3];
	if (!n_head) n_head = d_embed / d_head;
	x = N("norm", mlb_nn_groupnorm32(C, x));
	x = N("proj_in", mlb_nn_conv2d(C, x, d_embed, 1, 1, 1, 1, 0, 0, 1, 1, true));
	x = ggml_reshape_3d(C->cc, ggml_cont(C->cc, ggml_permute(C->cc, x, 1, 2, 0, 3)), d_embed, w * h, nb);
	for (int i = 0; i < depth; ++i) {
		snprintf(name, sizeof(name), "transf.%d", i);
		x = N(name, mlb_basic_transf(C, x, ctx, d_embed, d_embed, n_head));
	}
	x = ggml_reshape_4d(C->cc, ggml_cont(C->cc, ggml_permute(C->cc, x, 1, 0, 2, 3)), w, h, d_embed, nb);
	x = N("proj_out", mlb_nn_conv2d(C, x, (int)ch_in, 1, 1, 1, 1, 0, 0, 1, 1, true));
	return ggml_add(C->cc, x, skip);
}

MLTensor* mlb_unet_denoise(MLCtx* C, MLTensor* x, MLTensor* time, MLTensor* ctx, MLTensor* label, const UnetParams* P)
{
	char name[32];
	MLTensor* skips[32]; int n_skip = 0;
	mlctx_block_begin(C);

	/* timestep (+ SDXL label) embedding (unet.c:147-165) */
	MLTensor* emb = ggml_timestep_embedding(C->cc, time, P->n_ch, 10000);
	emb = ggml_silu_inplace(C->cc, N("time_embed.0", mlb_nn_linear(C, emb, P->n_te, true)));
	emb = N("time_embed.2", mlb_nn_linear(C, emb, P->n_te, true));
	if (P->ch_adm_in && label) {
		MLTensor* le = ggml_silu_inplace(C->cc, N("label_embed.0", mlb_nn_linear(C, label, P->n_te, true)));
		emb = ggml_add(C->cc, emb, N("label_embed.2", mlb_nn_linear(C, le, P->n_te, true)));
	}

	/* down path (unet.c:167-203) */
	x = N("in.conv", mlb_nn_conv2d(C, x, P->n_ch, 3, 3, 1, 1, 1, 1, 1, 1, true));
	skips[n_skip++] = x;
	int level = 0, blk = 0, ds = 1, ch = P->n_ch;
	for (; P->ch_mult[level]; ++level) {
		if (level) {
			ds *= 2;
			snprintf(name, sizeof(name), "in.%d.0", ++blk);
			x = N(name, mlb_downsample(C, x, ch, false));
			skips[n_skip++] = x;
		}
		for (int j = 0; j < P->n_res_blk; ++j) {
			ch = P->n_ch * P->ch_mult[level];
			snprintf(name, sizeof(name), "in.%d.0", ++blk);
			x = N(name, mlb_resnet(C, x, emb, ch));
			if (in_list(P->attn_res, ds)) {
				snprintf(name, sizeof(name), "in.%d.1", blk);
				x = N(name, mlb_spatial_transf(C, x, ctx, ch, P->d_head, P->n_head, P->transf_depth[level]));
			}
			skips[n_skip++] = x;
		}
	}
	level--;

	/* middle (unet.c:205-217) */
	x = N("mid.0", mlb_resnet(C, x, emb, ch));
	x = N("mid.1", mlb_spatial_transf(C, x, ctx, ch, P->d_head, P->n_head, P->transf_depth[level]));
	x = N("mid.2", mlb_resnet(C, x, emb, ch));

	/* up path (unet.c:219-258): concat skip (current first), resnet, [transformer], [upsample] */
	for (int ob = 0; level >= 0; --level) {
		for (int j = 0; j < P->n_res_blk + 1; ++j, ++ob) {
			GGML_ASSERT(n_skip > 0);
			x = ggml_concat(C->cc, x, skips[--n_skip], 2);
			int sub = 0;
			ch = P->n_ch * P->ch_mult[level];
			snprintf(name, sizeof(name), "out.%d.%d", ob, sub++);
			x = N(name, mlb_resnet(C, x, emb, ch));
			if (in_list(P->attn_res, ds)) {
				snprintf(name, sizeof(name), "out.%d.%d", ob, sub++);
				x = N(name, mlb_spatial_transf(C, x, ctx, ch, P->d_head, P->n_head, P->transf_depth[level]));
			}
			if (level && j == P->n_res_blk) {
				snprintf(name, sizeof(name), "out.%d.%d", ob, sub++);
				x = N(name, mlb_upsample(C, x, ch));
				ds /= 2;
			}
		}
	}
	GGML_ASSERT(n_skip == 0);
	x = ggml_silu_inplace(C->cc, N("out.norm", mlb_nn_groupnorm32(C, x)));
	return N("out.conv", mlb_nn_conv2d(C, x, P->n_ch_out, 3, 3, 1, 1, 1, 1, 1, 1, true));
}

/* Noise schedule of the trained model: scaled-linear betas 0.00085 -> 0.012, 1000 steps
 * (unet.c:283-304); log sigma table in float, accumulated in double. */
void unet_params_init(void)
{
	if (g_log_sigmas_sd[0] != 0) return;
	const int n = 1000;
	double b0 = sqrt(0.00085), b1 = sqrt(0.0120), step = (b1 - b0) / (n - 1), acp = 1.0;
	for (int i = 0; i < n; ++i) {
		double beta = b0 + step * i;
		acp *= 1.0 - beta * beta;
		g_log_sigmas_sd[i] = (float)log(sqrt((1 - acp) / acp));
	}
}

/* sigma -> fractional timestep (unet.c:305-328). The reference bisects for the FIRST table entry
 * >= log(sigma) and then extrapolates backwards along the following interval:
 *   t = idx + (ls - v[idx]) / (v[idx+1] - v[idx]),  clamped to n-1 at the top of the table.
 * That is not the textbook interpolation, but it is what feeds the timestep embedding. */
float unet_sigma_to_t(const UnetParams* P, float sigma)
{
	const float* v = P->log_sigmas; int n = P->n_step_train;
	float ls = (float)log(sigma);
	int lo = 0, hi = n;
	while (lo < hi) { int mid = (lo + hi) / 2; if (v[mid] - ls < 0) lo = mid + 1; else hi = mid; }
	int idx = lo;
	if (idx + 1 >= n) return (float)(n - 1);
	float v1 = v[idx], v2 = v[idx + 1];
	return idx + (ls - v1) / (v2 - v1);
}

float unet_t_to_sigma(const UnetParams* P, float t)
{
	const float* v = P->log_sigmas; int n = P->n_step_train;
	int ti = (int)t;
	if (ti < 0) ti = 0;
	if (ti > n - 1) ti = n - 1;
	float v1 = v[ti], v2 = ti + 1 < n ? v[ti + 1] : v1;
	float ls = v1 * (ti + 1 - t) + v2 * (t - ti);
	return (float)exp(ls);
}

int unet_denoise_init(UnetState* S, MLCtx* C, const UnetParams* P, unsigned lw, unsigned lh, int n_img, int n_rep)
{
	unet_params_init();
	bool reuse = C->prepared && S->ctx == C && S->par == P && S->lw == (int)lw && S->lh == (int)lh && S->n_img == n_img && S->n_rep == n_rep;
	S->nfe = 0;
	if (reuse) return 1;
	int nb = n_img * n_rep;
	C->c.n_tensor_max = 16384;
	mlctx_begin(C, "UNet");
	S->t_x = mlctx_input_new(C, "x", GGML_TYPE_F32, lw, lh, P->n_ch_in, nb);
	S->t_t = mlctx_input_new(C, "t", GGML_TYPE_F32, nb, 1, 1, 1);
	S->t_c = mlctx_input_new(C, "c", GGML_TYPE_F32, P->n_ctx, 77, nb, 1);
	S->t_l = P->ch_adm_in ? mlctx_input_new(C, "l", GGML_TYPE_F32, P->ch_adm_in, nb, 1, 1) : NULL;
	mlb_unet_denoise(C, S->t_x, S->t_t, S->t_c, S->t_l, P);
	C->c.tprefix = "unet";
	CHECK(mlctx_prep(C));
	S->ctx = C; S->par = P; S->lw = lw; S->lh = lh; S->n_img = n_img; S->n_rep = n_rep;
	return 1;
}

int unet_cond_set(UnetState* S, const float* cond, const float* label)
{
	ggml_backend_tensor_set(S->t_c, cond, 0, ggml_nbytes(S->t_c));
	if (S->t_l) {
		if (!label) FAIL(-1, "this model needs a label embedding");
		ggml_backend_tensor_set(S->t_l, label, 0, ggml_nbytes(S->t_l));
	}
	return 1;
}

int unet_denoise_run(UnetState* S, const float* x_dev, float sigma, const float** out_dev)
{
	const UnetParams* P = S->par;
	int nb = S->n_img * S->n_rep;
	int64_t n1 = (int64_t)S->lw * S->lh * P->n_ch_in * S->n_img;   /* elements of the latent batch */
	/* UNet input = x * c_in (unet.c:471-472), written straight into the graph input, once per CFG half */
	float c_in = 1 / sqrt(sigma * sigma + 1);   /* float sigma^2 + 1, double sqrt: the reference's expression */
	float* outs[2] = { (float*)S->t_x->data, (float*)S->t_x->data + n1 };
	const float* ins[1] = { x_dev };
	float coef[2] = { c_in, c_in };
	ggml_b200_lincomb(S->n_rep, outs, 1, ins, coef, n1);
	float tv[64];
	if (nb > 64) FAIL(-1, "batch too large");
	float t = unet_sigma_to_t(P, sigma);
	for (int i = 0; i < nb; ++i) tv[i] = t;
	ggml_backend_tensor_set(S->t_t, tv, 0, sizeof(float) * nb);
	CHECK(mlctx_compute(S->ctx));
	S->nfe += S->n_rep;
	*out_dev = (const float*)S->ctx->result->data;
	return 1;
}
