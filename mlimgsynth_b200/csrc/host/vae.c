#include "vae.h"
#include "mlblock_nn.h"

#define N(name, x) mlctx_tensor_add(C, (name), (x))

const VaeParams g_vae_sd1  = { 3, 4, 128, 4, 2, {1, 2, 4, 4}, 4, 8, 0.18215f };
const VaeParams g_vae_sdxl = { 3, 4, 128, 4, 2, {1, 2, 4, 4}, 4, 8, 0.13025f };
const SdTaeParams g_sdtae_sd1 = { 3, 64, 4, 3 };

/* single-head spatial self-attention of the VAE middle block (vae.c:46-74) */
static MLTensor* mlb_attn_2d_self(MLCtx* C, MLTensor* x)
{
	MLTensor* skip = x;
	mlctx_block_begin(C);
	x = N("norm", mlb_nn_groupnorm32(C, x));
	int64_t w = x->ne[0], h = x->ne[1], c = x->ne[2], n = x->ne[3];
	MLTensor* q = N("q", mlb_nn_conv2d(C, x, (int)c, 1, 1, 1, 1, 0, 0, 1, 1, true));
	q = ggml_reshape_3d(C->cc, ggml_cont(C->cc, ggml_permute(C->cc, q, 1, 2, 0, 3)), c, h * w, n);
	MLTensor* k = N("k", mlb_nn_conv2d(C, x, (int)c, 1, 1, 1, 1, 0, 0, 1, 1, true));
	k = ggml_reshape_3d(C->cc, ggml_cont(C->cc, ggml_permute(C->cc, k, 1, 2, 0, 3)), c, h * w, n);
	MLTensor* v = N("v", mlb_nn_conv2d(C, x, (int)c, 1, 1, 1, 1, 0, 0, 1, 1, true));
	v = ggml_reshape_3d(C->cc, v, h * w, c, n);
	x = mlb_attention(C, q, k, v, false);
	x = ggml_reshape_4d(C->cc, ggml_cont(C->cc, ggml_permute(C->cc, x, 1, 0, 2, 3)), w, h, c, n);
	x = N("proj_out", mlb_nn_conv2d(C, x, (int)c, 1, 1, 1, 1, 0, 0, 1, 1, true));
	return ggml_add(C->cc, x, skip);
}

static MLTensor* kl_mid(MLCtx* C, MLTensor* x, int ch)
{
	x = N("mid.block_1", mlb_resnet(C, x, NULL, ch));
	x = N("mid.attn_1", mlb_attn_2d_self(C, x));
	return N("mid.block_2", mlb_resnet(C, x, NULL, ch));
}
static MLTensor* kl_out(MLCtx* C, MLTensor* x, int ch_out)
{
	x = ggml_silu_inplace(C->cc, N("norm_out", mlb_nn_groupnorm32(C, x)));
	return N("conv_out", mlb_nn_conv2d(C, x, ch_out, 3, 3, 1, 1, 1, 1, 1, 1, true));
}

static MLTensor* mlb_kl_encoder(MLCtx* C, MLTensor* x, int ch_out, const VaeParams* P)
{
	char name[48];
	mlctx_block_begin(C);
	x = N("conv_in", mlb_nn_conv2d(C, x, P->ch, 3, 3, 1, 1, 1, 1, 1, 1, true));
	int cb = P->ch;
	for (int i = 0; i < P->n_res; ++i) {
		for (int j = 0; j < P->n_res_blk; ++j) {
			snprintf(name, sizeof(name), "down.%d.block.%d", i, j);
			cb = P->ch * P->ch_mult[i];
			x = N(name, mlb_resnet(C, x, NULL, cb));
		}
		if (i + 1 != P->n_res) {
			snprintf(name, sizeof(name), "down.%d.downsample", i);
			x = N(name, mlb_downsample(C, x, cb, true));
		}
	}
	return kl_out(C, kl_mid(C, x, cb), ch_out);
}

static MLTensor* mlb_kl_decoder(MLCtx* C, MLTensor* x, int ch_out, const VaeParams* P)
{
	char name[48];
	mlctx_block_begin(C);
	int cb = P->ch * P->ch_mult[P->n_res - 1];
	x = N("conv_in", mlb_nn_conv2d(C, x, cb, 3, 3, 1, 1, 1, 1, 1, 1, true));
	x = kl_mid(C, x, cb);
	for (int i = P->n_res - 1; i >= 0; --i) {
		for (int j = 0; j < P->n_res_blk + 1; ++j) {
			snprintf(name, sizeof(name), "up.%d.block.%d", i, j);
			cb = P->ch * P->ch_mult[i];
			x = N(name, mlb_resnet(C, x, NULL, cb));
		}
		if (i) {
			snprintf(name, sizeof(name), "up.%d.upsample", i);
			x = N(name, mlb_upsample(C, x, cb));
		}
	}
	return kl_out(C, x, ch_out);
}

MLTensor* mlb_sdvae_encoder(MLCtx* C, MLTensor* x, const VaeParams* P)
{
	x = N("encoder", mlb_kl_encoder(C, x, P->ch_z * 2, P));
	return N("quant_conv", mlb_nn_conv2d(C, x, P->ch_z * 2, 1, 1, 1, 1, 0, 0, 1, 1, true));
}

MLTensor* mlb_sdvae_decoder(MLCtx* C, MLTensor* x, const VaeParams* P)
{
	x = ggml_scale(C->cc, x, 1 / P->scale_factor);
	x = N("post_quant_conv", mlb_nn_conv2d(C, x, P->d_embed, 1, 1, 1, 1, 0, 0, 1, 1, true));
	return N("decoder", mlb_kl_decoder(C, x, P->ch_x, P));
}

/* ---- TAESD (tae.c:24-92) ---- */
static MLTensor* tae_block(MLCtx* C, MLTensor* x, int ch)
{
	MLTensor* skip = x;
	mlctx_block_begin(C);
	x = ggml_relu_inplace(C->cc, N("conv.0", mlb_nn_conv2d(C, x, ch, 3, 3, 1, 1, 1, 1, 1, 1, true)));
	x = ggml_relu_inplace(C->cc, N("conv.2", mlb_nn_conv2d(C, x, ch, 3, 3, 1, 1, 1, 1, 1, 1, true)));
	x = N("conv.4", mlb_nn_conv2d(C, x, ch, 3, 3, 1, 1, 1, 1, 1, 1, true));
	return ggml_relu_inplace(C->cc, ggml_add(C->cc, x, skip));
}
#define IDX(i) (snprintf(name, sizeof(name), "%d", (i)), name)

MLTensor* mlb_sdtae_encoder(MLCtx* C, MLTensor* x, const SdTaeParams* P)
{
	char name[16]; int b = 0;
	mlctx_block_begin(C);
	x = N(IDX(b), mlb_nn_conv2d(C, x, P->ch_inner, 3, 3, 1, 1, 1, 1, 1, 1, true)); b++;
	x = N(IDX(b), tae_block(C, x, P->ch_inner)); b++;
	for (int j = 0; j < 3; ++j) {
		x = N(IDX(b), mlb_nn_conv2d(C, x, P->ch_inner, 3, 3, 2, 2, 1, 1, 1, 1, false)); b++;
		for (int i = 0; i < P->n_blk; ++i) { x = N(IDX(b), tae_block(C, x, P->ch_inner)); b++; }
	}
	return N(IDX(b), mlb_nn_conv2d(C, x, P->ch_z, 3, 3, 1, 1, 1, 1, 1, 1, true));
}

MLTensor* mlb_sdtae_decoder(MLCtx* C, MLTensor* x, const SdTaeParams* P)
{
	char name[16]; int b = 0;
	mlctx_block_begin(C);
	x = ggml_scale(C->cc, ggml_tanh_inplace(C->cc, ggml_scale(C->cc, x, 1.0f / 3.0f)), 3.0f);   /* 3 tanh(x/3) */
	x = ggml_relu_inplace(C->cc, N(IDX(b), mlb_nn_conv2d(C, x, P->ch_inner, 3, 3, 1, 1, 1, 1, 1, 1, true))); b += 2;
	for (int j = 0; j < 3; ++j) {
		for (int i = 0; i < P->n_blk; ++i) { x = N(IDX(b), tae_block(C, x, P->ch_inner)); b++; }
		x = ggml_upscale(C->cc, x, 2, GGML_SCALE_MODE_NEAREST); b++;
		x = N(IDX(b), mlb_nn_conv2d(C, x, P->ch_inner, 3, 3, 1, 1, 1, 1, 1, 1, false)); b++;
	}
	x = N(IDX(b), tae_block(C, x, P->ch_inner)); b++;
	return N(IDX(b), mlb_nn_conv2d(C, x, P->ch_x, 3, 3, 1, 1, 1, 1, 1, 1, true));
}

/* ---- runners ---- */
static int codec_prepare(CodecState* S, MLCtx* C, int kind, int n0, int n1, int nb, const void* P)
{
	if (C->prepared && S->ctx == C && S->kind == kind && S->n0 == n0 && S->n1 == n1 && S->nb == nb) return 1;
	static const char* names[] = { "", "VAE decode", "VAE encode", "TAE decode", "TAE encode" };
	mlctx_begin(C, names[kind]);
	int cin = (kind == CODEC_VAE_DEC || kind == CODEC_TAE_DEC) ? 4 : 3;
	S->t_in = mlctx_input_new(C, "in", GGML_TYPE_F32, n0, n1, cin, nb);
	switch (kind) {
	case CODEC_VAE_DEC: S->t_out = mlb_sdvae_decoder(C, S->t_in, P); C->c.tprefix = "vae"; break;
	case CODEC_VAE_ENC: S->t_out = mlb_sdvae_encoder(C, S->t_in, P); C->c.tprefix = "vae"; break;
	case CODEC_TAE_DEC: S->t_out = mlb_sdtae_decoder(C, S->t_in, P); mlctx_tensor_add(C, "decoder.layers", S->t_out); C->c.tprefix = "tae"; break;
	case CODEC_TAE_ENC: S->t_out = mlb_sdtae_encoder(C, S->t_in, P); mlctx_tensor_add(C, "encoder.layers", S->t_out); C->c.tprefix = "tae"; break;
	}
	CHECK(mlctx_prep(C));
	S->ctx = C; S->kind = kind; S->n0 = n0; S->n1 = n1; S->nb = nb;
	return 1;
}

/* plane-wise region copy between a full tensor and a tile, on the device (ltensor_copy_slice2 role) */
static void copy_region(float* dst, int dw, int dh, int dx, int dy, const float* src, int sw, int sh, int sx, int sy,
	int w, int h, int planes)
{
	if (w <= 0 || h <= 0) return;
	for (int p = 0; p < planes; ++p)
		ggml_b200_copy2d(dst + ((size_t)p * dh + dy) * dw + dx, (size_t)dw * sizeof(float),
			src + ((size_t)p * sh + sy) * sw + sx, (size_t)sw * sizeof(float), (size_t)w * sizeof(float), (size_t)h);
}

/* Tiled run of a codec graph whose output is `up/down` times the input size. Geometry as the
 * reference (vae.c:229-260 encode, :331-391 decode): tile = requested size + 2k clipped to the
 * tensor, step = tile - 2k, the last tile is shifted back inside the tensor; tiles are visited in
 * row-major order and later tiles overwrite earlier ones (no blending); of each tile the region
 * [d, d + n - k) is kept, with d = k except at the left/top border. */
static int run_tiled(CodecState* S, const float* in_dev, int iw, int ih, int cin, float* out_dev, int cout,
	int n0, int n1, int k, int up, int down, float in_mul, float in_post)
{
	MLCtx* C = S->ctx;
	const int ow = iw * up / down, oh = ih * up / down, tw_o = n0 * up / down, th_o = n1 * up / down;
	float* tin = (float*)S->t_in->data;
	const float* tout = (const float*)S->t_out->data;
	const bool single = (n0 == iw && n1 == ih);
	const int step0 = single ? n0 : n0 - 2 * k, step1 = single ? n1 : n1 - 2 * k;
	const int nt0 = (iw + step0 - 1) / step0, nt1 = (ih + step1 - 1) / step1;
	for (int t1 = 0; t1 < nt1; ++t1) {
		int i1 = t1 * step1; if (i1 > ih - n1) i1 = ih - n1;
		for (int t0 = 0; t0 < nt0; ++t0) {
			int i0 = t0 * step0; if (i0 > iw - n0) i0 = iw - n0;
			copy_region(tin, n0, n1, 0, 0, in_dev, iw, ih, i0, i1, n0, n1, cin);
			if (in_mul != 1 || in_post != 0) ggml_b200_affine(tin, tin, 0, in_mul, in_post, (int64_t)n0 * n1 * cin);
			CHECK(mlctx_compute(C));
			if (single) { copy_region(out_dev, ow, oh, 0, 0, tout, tw_o, th_o, 0, 0, tw_o, th_o, cout); continue; }
			int d0 = i0 ? k : 0, d1 = i1 ? k : 0;
			/* a dim that one tile covers entirely is copied whole (the reference leaves its last k rows/columns
			 * unwritten in that case, vae.c:381-384 with n == full size) */
			int c0 = n0 == iw ? n0 : n0 - k, c1 = n1 == ih ? n1 : n1 - k;
			copy_region(out_dev, ow, oh, (i0 + d0) * up / down, (i1 + d1) * up / down, tout, tw_o, th_o,
				d0 * up / down, d1 * up / down, c0 * up / down, c1 * up / down, cout);
		}
	}
	return 1;
}

static int tile_extent(int tile_px, int unit, int k, int full)
{
	if (tile_px <= 0) return full;
	tile_px = (tile_px + 63) / 64 * 64;
	int n = tile_px / unit + 2 * k;
	return n < full ? n : full;
}

int sdvae_decode(CodecState* S, MLCtx* C, const VaeParams* P, const float* latent_dev, int lw, int lh, float* image_dev, int tile_px)
{
	const int f = P->f_down, k = 8;
	int n0 = tile_extent(tile_px, f, k, lw), n1 = tile_extent(tile_px, f, k, lh);
	CHECK(codec_prepare(S, C, CODEC_VAE_DEC, n0, n1, 1, P));
	CHECK(run_tiled(S, latent_dev, lw, lh, 4, image_dev, 3, n0, n1, k, f, 1, 1, 0));
	int64_t n = (int64_t)lw * f * lh * f * 3;
	ggml_b200_affine(image_dev, image_dev, 1, 0.5f, 0, n);   /* (x+1)/2: [-1,1] -> [0,1] (vae.h:43-47) */
	return 1;
}

/* ---- tiled decode split across workers (VAE tiles are independent given the fixed tile graph, vae.c:343-346) ----
 * A worker (GPU rank) decodes the tiles  first, first + stride, ...  of the reference's row-major tile list into consecutive
 * slots of `tiles_dev`; after the slots of all workers are gathered, the merge pastes them in the reference's visiting order
 * so that overlaps resolve exactly as in the serial loop (later tiles overwrite earlier ones, vae.c:365-387). */
int sdvae_tile_plan(const VaeParams* P, int lw, int lh, int tile_px, VaeTilePlan* T)
{
	const int f = P->f_down, k = 8;
	T->k = k; T->f = f;
	T->n0 = tile_extent(tile_px, f, k, lw); T->n1 = tile_extent(tile_px, f, k, lh);
	const bool single = (T->n0 == lw && T->n1 == lh);
	T->step0 = single ? T->n0 : T->n0 - 2 * k; T->step1 = single ? T->n1 : T->n1 - 2 * k;
	T->nt0 = (lw + T->step0 - 1) / T->step0; T->nt1 = (lh + T->step1 - 1) / T->step1;
	T->tile_elems = (size_t)T->n0 * f * T->n1 * f * P->ch_x;
	return T->nt0 * T->nt1;
}

int sdvae_decode_tiles(CodecState* S, MLCtx* C, const VaeParams* P, const float* latent_dev, int lw, int lh, int tile_px,
	int first, int stride, float* tiles_dev)
{
	VaeTilePlan T;
	const int nt = sdvae_tile_plan(P, lw, lh, tile_px, &T);
	if (first < 0 || stride < 1) FAIL(-1, "invalid tile assignment %d/%d", first, stride);
	CHECK(codec_prepare(S, C, CODEC_VAE_DEC, T.n0, T.n1, 1, P));
	float* tin = (float*)S->t_in->data;
	for (int t = first, slot = 0; t < nt; t += stride, ++slot) {
		int t1 = t / T.nt0, t0 = t % T.nt0;
		int i1 = t1 * T.step1; if (i1 > lh - T.n1) i1 = lh - T.n1;
		int i0 = t0 * T.step0; if (i0 > lw - T.n0) i0 = lw - T.n0;
		copy_region(tin, T.n0, T.n1, 0, 0, latent_dev, lw, lh, i0, i1, T.n0, T.n1, 4);
		CHECK(mlctx_compute(C));
		ggml_b200_copy(tiles_dev + T.tile_elems * slot, S->t_out->data, T.tile_elems * sizeof(float));
	}
	return nt;
}

int sdvae_merge_tiles(const VaeParams* P, int lw, int lh, int tile_px, const float* gathered_dev, int world, int slots_per_worker, float* image_dev)
{
	VaeTilePlan T;
	const int nt = sdvae_tile_plan(P, lw, lh, tile_px, &T);
	const int f = T.f, k = T.k, ow = lw * f, oh = lh * f, tw_o = T.n0 * f, th_o = T.n1 * f;
	if (world < 1 || slots_per_worker * world < nt) FAIL(-1, "tile buffer too small: %d x %d slots for %d tiles", world, slots_per_worker, nt);
	/* one launch: every output pixel takes the LAST tile of the reference's row-major list (t1 outer, t0 inner) whose kept region
	   covers it -- exactly what pasting the tiles one after the other produces (vae.c:365-387) -- and (x+1)/2 is applied on the way
	   (vae.h:43-47). Kept region of a tile, in latent pixels: [d, d + n - k) with d = k except at the origin; a dim that one tile covers
	   entirely is copied whole (same rule as run_tiled). */
	(void)nt;
	ggml_b200_tile_merge(image_dev, gathered_dev, ow, oh, P->ch_x, T.nt0, T.nt1, tw_o, th_o, T.step0 * f, T.step1 * f, k * f,
		T.n0 == lw, T.n1 == lh, world, slots_per_worker, 1.0f, 0.5f);
	return nt;
}

/* Untiled decode of nb latents in ONE graph run: the images of a batch are independent, so the decoder is
 * built with a batch dimension -- the 64x64 / 128x128 levels get GEMM rows from all images and every launch is shared.
 * Returns 0 (nothing done) when the requested tile size makes the decode tiled: the caller then decodes image by image
 * in the reference's tile order. */
int sdvae_decode_batch(CodecState* S, MLCtx* C, const VaeParams* P, const float* latent_dev, int lw, int lh, int nb, float* image_dev, int tile_px)
{
	const int f = P->f_down, k = 8;
	if (tile_extent(tile_px, f, k, lw) != lw || tile_extent(tile_px, f, k, lh) != lh || nb < 2) return 0;
	CHECK(codec_prepare(S, C, CODEC_VAE_DEC, lw, lh, nb, P));
	const size_t n_in = (size_t)lw * lh * 4 * nb, n_out = (size_t)lw * f * lh * f * 3 * nb;
	ggml_b200_copy(S->t_in->data, latent_dev, n_in * sizeof(float));   /* same [w,h,c,n] layout */
	CHECK(mlctx_compute(C));
	ggml_b200_copy(image_dev, S->t_out->data, n_out * sizeof(float));
	ggml_b200_affine(image_dev, image_dev, 1, 0.5f, 0, (int64_t)n_out);   /* (x+1)/2: [-1,1] -> [0,1] (vae.h:43-47) */
	return 1;
}

int sdvae_encode(CodecState* S, MLCtx* C, const VaeParams* P, const float* image_dev, int w, int h, float* moments_dev, int tile_px)
{
	const int f = P->f_down, k = f * 8;
	if (w % f || h % f) FAIL(-1, "invalid input image size %dx%d", w, h);
	int n0 = tile_extent(tile_px, 1, k, w), n1 = tile_extent(tile_px, 1, k, h);
	CHECK(codec_prepare(S, C, CODEC_VAE_ENC, n0, n1, 1, P));
	/* [0,1] -> [-1,1] on the tile (vae.h:36-41), then encode */
	return run_tiled(S, image_dev, w, h, 3, moments_dev, 8, n0, n1, k, 1, f, 2, -1);
}

int sdtae_decode(CodecState* S, MLCtx* C, const SdTaeParams* P, const float* latent_dev, int lw, int lh, float* image_dev)
{
	CHECK(codec_prepare(S, C, CODEC_TAE_DEC, lw, lh, 1, P));
	return run_tiled(S, latent_dev, lw, lh, 4, image_dev, 3, lw, lh, 0, 8, 1, 1, 0);   /* output already in [0,1] */
}

int sdtae_encode(CodecState* S, MLCtx* C, const SdTaeParams* P, const float* image_dev, int w, int h, float* latent_dev)
{
	if (w % 8 || h % 8) FAIL(-1, "invalid input image size %dx%d", w, h);
	CHECK(codec_prepare(S, C, CODEC_TAE_ENC, w, h, 1, P));
	return run_tiled(S, image_dev, w, h, 3, latent_dev, 4, w, h, 0, 1, 8, 1, 0);
}
