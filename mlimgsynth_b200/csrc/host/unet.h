/* unet.h -- SD 1.x / 2.x / XL denoiser: graph builder and device-resident evaluation.
 * Mirrors the reference's unet.h (UnetParams, g_unet_*, mlb_unet_denoise, unet_sigma_to_t,
 * unet_t_to_sigma, unet_denoise_init/run). B200 differences: the graph carries a batch
 * (images x CFG halves) instead of N=1 (unet.c:351-355), the latent never leaves the device, and
 * the c_in / v-prediction / CFG arithmetic is fused into the sampler kernels (ggml-b200.h). */
#pragma once
#include "mlblock.h"

typedef struct UnetParams {
	int n_ch_in, n_ch_out, n_res_blk;
	int attn_res[4], ch_mult[5], transf_depth[5];
	int n_te, n_head, d_head, n_ctx, n_ch, ch_adm_in;
	unsigned clip_norm:1, cond_label:1, uncond_empty_zero:1, vparam:1;
	int n_step_train;
	float sigma_min, sigma_max;
	float* log_sigmas;
} UnetParams;

extern const UnetParams g_unet_sd1, g_unet_sd2, g_unet_sdxl;

MLTensor* mlb_unet_denoise(MLCtx* C, MLTensor* x, MLTensor* time, MLTensor* c, MLTensor* label, const UnetParams* P);

void  unet_params_init(void);
float unet_sigma_to_t(const UnetParams* P, float sigma);
float unet_t_to_sigma(const UnetParams* P, float t);

typedef struct UnetState {
	MLCtx* ctx;
	const UnetParams* par;
	unsigned nfe;
	int lw, lh, n_img, n_rep;      /* latent size, images, graph batch = n_img * n_rep (n_rep: 1, or 2 with CFG) */
	MLTensor *t_x, *t_t, *t_c, *t_l;
} UnetState;

/* Builds (or reuses, when the shape is unchanged) the batched UNet graph and uploads its weights. */
int unet_denoise_init(UnetState* S, MLCtx* C, const UnetParams* P, unsigned lw, unsigned lh, int n_img, int n_rep);
/* cond rows: [n_ctx, 77, n_img*n_rep] host data, label [ch_adm_in, n_img*n_rep] (or NULL); uploaded once per generation */
int unet_cond_set(UnetState* S, const float* cond, const float* label);
/* One batched evaluation at noise level sigma: reads the device latent x [lw,lh,4,n_img], writes the
 * graph output (device, [lw,lh,4,n_img*n_rep]); returns its device pointer. c_in is applied on device. */
int unet_denoise_run(UnetState* S, const float* x_dev, float sigma, const float** out_dev);
