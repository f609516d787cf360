/* vae.h -- KL autoencoder of SD (decoder + encoder graphs, tiling, posterior sampling) and the
 * tiny autoencoder TAESD. Mirrors the reference's vae.h / tae.h; graphs are cached per tile shape,
 * latents and images move device-to-device, pre/post scaling runs in the pack/unpack kernels. */
#pragma once
#include "mlblock.h"

typedef struct VaeParams { int ch_x, ch_z, ch, n_res, n_res_blk, ch_mult[5], d_embed, f_down; float scale_factor; } VaeParams;
extern const VaeParams g_vae_sd1, g_vae_sdxl;
typedef struct SdTaeParams { int ch_x, ch_inner, ch_z, n_blk; } SdTaeParams;
extern const SdTaeParams g_sdtae_sd1;

MLTensor* mlb_sdvae_encoder(MLCtx* C, MLTensor* x, const VaeParams* P);
MLTensor* mlb_sdvae_decoder(MLCtx* C, MLTensor* x, const VaeParams* P);
MLTensor* mlb_sdtae_encoder(MLCtx* C, MLTensor* x, const SdTaeParams* P);
MLTensor* mlb_sdtae_decoder(MLCtx* C, MLTensor* x, const SdTaeParams* P);

/* One cached codec graph (input tile shape -> output tile shape). */
typedef struct CodecState { MLCtx* ctx; int kind, n0, n1, nb; MLTensor *t_in, *t_out; } CodecState;
enum { CODEC_VAE_DEC = 1, CODEC_VAE_ENC = 2, CODEC_TAE_DEC = 3, CODEC_TAE_ENC = 4 };

/* latent_dev [lw,lh,4] (device f32, SD-scaled) -> image_dev [8lw,8lh,3] (device f32, in [0,1]).
 * tile_px > 0 decodes overlapping tiles in the reference's order and geometry (vae.c:318-410). */
int sdvae_decode(CodecState* S, MLCtx* C, const VaeParams* P, const float* latent_dev, int lw, int lh, float* image_dev, int tile_px);
/* nb latents [lw,lh,4,nb] -> images [8lw,8lh,3,nb] in one run of a batched decoder graph; returns 0 if tile_px asks for tiling */
int sdvae_decode_batch(CodecState* S, MLCtx* C, const VaeParams* P, const float* latent_dev, int lw, int lh, int nb, float* image_dev, int tile_px);
/* Tiled decode split across workers: plan (tile list geometry of vae.c:331-368), decode of the tiles first, first+stride, ...
 * into consecutive slots of tiles_dev (raw decoder output, tile_elems floats each), and the merge of the gathered slots
 * (tile t lives at slot (t % world) * slots_per_worker + t / world) in the reference's row-major order + (x+1)/2. */
typedef struct VaeTilePlan { int n0, n1, nt0, nt1, step0, step1, k, f; size_t tile_elems; } VaeTilePlan;
int sdvae_tile_plan(const VaeParams* P, int lw, int lh, int tile_px, VaeTilePlan* plan);
int sdvae_decode_tiles(CodecState* S, MLCtx* C, const VaeParams* P, const float* latent_dev, int lw, int lh, int tile_px,
	int first, int stride, float* tiles_dev);
int sdvae_merge_tiles(const VaeParams* P, int lw, int lh, int tile_px, const float* gathered_dev, int world, int slots_per_worker, float* image_dev);
/* image_dev [w,h,3] in [0,1] -> moments_dev [w/8,h/8,8] (mean | logvar), tiled like vae.c:222-316 */
int sdvae_encode(CodecState* S, MLCtx* C, const VaeParams* P, const float* image_dev, int w, int h, float* moments_dev, int tile_px);
int sdtae_decode(CodecState* S, MLCtx* C, const SdTaeParams* P, const float* latent_dev, int lw, int lh, float* image_dev);
int sdtae_encode(CodecState* S, MLCtx* C, const SdTaeParams* P, const float* image_dev, int w, int h, float* latent_dev);
