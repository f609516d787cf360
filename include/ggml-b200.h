/* ggml-b200.h -- engine extras beyond the ggml-shaped ABI: device-resident sampler state.
 *
 * The reference keeps the latent on the host and, per UNet evaluation, loops over it on the CPU
 * (c_in scaling unet.c:471-472, v-prediction mix unet.c:490-494, CFG combine mlimgsynth.c:1583,
 * solver updates solvers.c:86,107-115,149-163,221-227,280-286, noise add / inpaint mask
 * sampling.c:98-117) with 3-4 uploads and one download around every graph run (unet.c:375-384).
 * Every one of those loops is a linear combination of a few same-shaped f32 tensors with scalar
 * coefficients the host computes exactly as the reference does, so the B200 host layer keeps the
 * latent in HBM and issues ONE fused, vectorised kernel per solver stage on the engine's stream.
 * All calls are asynchronous w.r.t. the host except ggml_b200_download / _nonfinite_check.
 */
#ifndef GGML_B200_EXTRAS_H
#define GGML_B200_EXTRAS_H
#include "ggml-backend.h"
#ifdef __cplusplus
extern "C" {
#endif

#define GGML_B200_LINCOMB_MAX_IN  6
#define GGML_B200_LINCOMB_MAX_OUT 3

/* device memory (f32 working tensors of the sampler) */
GGML_API void* ggml_b200_malloc(size_t bytes);
GGML_API void  ggml_b200_free(void* dev);
GGML_API void  ggml_b200_upload(void* dev, const void* host, size_t bytes);      /* async, host buffer reusable on return */
GGML_API void  ggml_b200_download(void* host, const void* dev, size_t bytes);    /* synchronises the engine stream */
GGML_API void* ggml_b200_host_malloc(size_t bytes);                               /* page-locked host memory */
GGML_API void  ggml_b200_host_free(void* p);
GGML_API void  ggml_b200_copy(void* dst, const void* src, size_t bytes);         /* device to device */
GGML_API void  ggml_b200_memset(void* dev, int value, size_t bytes);
/* rows x width_bytes region copy with pitches (device to device; tile gather/scatter of the VAE tiling) */
GGML_API void  ggml_b200_copy2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width_bytes, size_t rows);
/* out = (in + pre_add) * mul + post_add  (VAE pre/post scaling vae.h:36-47, exact in that order) */
GGML_API void  ggml_b200_affine(float* out, const float* in, float pre_add, float mul, float post_add, int64_t n);
GGML_API void  ggml_b200_synchronize(void);

/* out[j][e] = sum_i coef[j*n_in + i] * in[i][e],  e < n;  n_out <= 3, n_in <= 6.
 * All inputs of element e are read before any output is written: outputs may alias inputs. */
GGML_API void ggml_b200_lincomb(int n_out, float* const* outs, int n_in, const float* const* ins,
	const float* coef, int64_t n);

/* Inpainting blend (sampling.c:98-110): x[p,c] = x0[p,c]*m[p] + x[p,c]*(1-m[p]); planes of n_pix, n_planes planes. */
GGML_API void ggml_b200_mask_blend(float* x, const float* x0, const float* mask, int64_t n_pix, int64_t n_planes);

/* VAE posterior sample (vae.c:197-220): out = (mean + exp(0.5*clamp(logvar,-30,20)) * noise) * scale */
GGML_API void ggml_b200_vae_sample(float* out, const float* mean, const float* logvar, const float* noise,
	float scale, int64_t n);

/* Planar f32 image [W,H,C] -> interleaved RGB8: v = x*mul+add; clamp(v*255, 0, 255) truncated
 * (vae.h:43-47 post-scaling + mlimgsynth.c:112-129). */
GGML_API void ggml_b200_pack_rgb8(uint8_t* out_hwc, const float* in_chw, int w, int h, int c, float mul, float add);
/* interleaved u8 -> planar f32 * (1/255) (mlimgsynth.c:131-151), optional mul/add pre-scaling */
GGML_API void ggml_b200_unpack_u8(float* out_chw, const uint8_t* in_hwc, int w, int h, int c_in, int c_first, int c_count,
	float mul, float add);
/* box average fw x fh (localtensor.c:161-194, mask downsample 8x8) */
GGML_API void ggml_b200_box_downsample(float* out, const float* in, int w, int h, int planes, int fw, int fh);

/* LoRA merge on device (lora.c:46-78): W[n1][n0] (f16) <- f16( f32(W) + scale * (up[n1][r] . down[r][n0]) ),
 * operands f16, products accumulated in f32, ONE f16 rounding per merge (same as the reference). */
/* Ordered merge of decoded tiles (all geometry in OUTPUT pixels): tile t = t1 * nt0 + t0 starts at min(t0 * step0, ow - tw) and keeps
 * [margin, margin + tw - margin) of its extent (margin 0 at the origin; the whole extent when one tile covers the axis); every pixel takes the last
 * tile of the row-major list that covers it; tile t lives at slot (t % world) * slots + t / world of `tiles`; out = (v + pre_add) * mul. */
GGML_API void ggml_b200_tile_merge(float* out, const float* tiles, int ow, int oh, int planes, int nt0, int nt1, int tw, int th, int step0, int step1,
	int keep_margin, int full0, int full1, int world, int slots, float pre_add, float mul);
GGML_API void ggml_b200_lora_merge_f16(void* w_dev, const void* down_dev, const void* up_dev, int64_t n0, int64_t n1, int r, float scale);
GGML_API void ggml_b200_lora_merge_f32(void* w_dev, const void* down_dev, const void* up_dev, int64_t n0, int64_t n1, int r, float scale);

/* Non-finite guard (unet.c:487): accumulates into a device flag; _check downloads and clears it. */
GGML_API void ggml_b200_nonfinite_accumulate(const float* x, int64_t n);
GGML_API int  ggml_b200_nonfinite_check(void);

/* Per-kernel-family profile: while enabled, graphs run eagerly with CUDA events around every step
 * (measurement aid for bench.py's roofline; not used in timed regions). kind: 7 simt gemm,
 * 8 groupnorm, 9 layernorm, 10 geglu, 13 attention, 14 tcgen05 gemm, 15 tcgen05 conv3x3, 0 copy. */
GGML_API void ggml_b200_profile_enable(int on);
GGML_API int  ggml_b200_profile_get(int kind, double* ms, double* flops, double* bytes, uint64_t* launches);
/* GPU time between two points of the engine stream (CUDA events): start, then stop returns ms (synchronises). */
GGML_API void   ggml_b200_timer_start(void);
GGML_API double ggml_b200_timer_stop(void);

/* counters: kernels launched by this library, graph replays, bytes moved across PCIe */
struct ggml_b200_stats { uint64_t kernel_launches, graph_launches, plans_built, h2d_bytes, d2h_bytes; };
GGML_API const struct ggml_b200_stats* ggml_b200_get_stats(void);
/* Steps of the most recently planned graph: key = a step kind ("GEMM_TC", "CONV_TC", "ATTENTION", "GROUPNORM", "LAYERNORM",
 * "COPY", ...), "name:<step name>" (e.g. "name:linear_fused", "name:linear_geglu"), "steps" or "preps" (totals). */
GGML_API int ggml_b200_last_plan_count(const char* key);

#ifdef __cplusplus
}
#endif
#endif
