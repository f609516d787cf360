// kernels.h -- launch interface between the planner and the sm_100a kernels.
#pragma once
#include "engine.h"

namespace b200 {

// ---- generic strided kernels (kernels_elem.cu): correct for any layout, used for graph
// inputs/outputs, rare ops and small tensors. Hot tensors use the dedicated kernels below.
void k_copy(cudaStream_t s, const View& dst, const View& src);
void k_binary(cudaStream_t s, BinOp op, const View& dst, const View& a, const View& b);
void k_unary(cudaStream_t s, UnaryOp op, float param, const View& dst, const View& src);
void k_upscale(cudaStream_t s, const View& dst, const View& src);
void k_spin(cudaStream_t s, int microseconds);      // profiling aid: keeps the stream busy while the next launches are enqueued
void k_softmax_rows(cudaStream_t s, const View& dst, const View& src, bool causal, int n_past);
// rows of f32 scores -> f16 probabilities softmax(scale * s) (wide-head attention as two tensor-core GEMMs)
bool k_softmax_f32_f16_supported(int64_t cols);
void k_softmax_f32_f16(cudaStream_t s, const float* scores, __half* probs, int64_t rows, int64_t cols, int64_t ld_s, int64_t ld_p, float scale);
void k_get_rows(cudaStream_t s, const View& dst, const View& table, const View& ids);
void k_timestep_embedding(cudaStream_t s, const View& dst, const View& ts, int dim, int max_period);
void k_gemm_simt(cudaStream_t s, const View& c, const View& a, const View& b, bool round_b_f16);

// ---- memory-bound fused kernels on the engine's native layout (rows of channels, f16)
// GroupNorm(+affine)(+SiLU): src/dst [W,H,C,N] with channel stride 1 (channels-last).
// stats: 2*N*groups doubles, zeroed by the caller before launch.
void k_groupnorm(cudaStream_t s, const View& dst, const View& src, const float* gamma, const float* beta,
	int groups, float eps, bool silu, unsigned long long* stats, bool stats_ready = false);   // stats_ready: the producer's epilogue filled them
// stats slot: [N][groups][4] 64-bit words (fixed-point sum / sum of squares, integer atomics: order-independent), zeroed per run
// LayerNorm(+affine) over dim 0 (stride 1), one warp per row.
void k_layernorm(cudaStream_t s, const View& dst, const View& src, const float* gamma, const float* beta, float eps);
// GEGLU gate: dst[j,m] = h[j,m] * gelu(h[d+j,m]); h rows have 2d channels.
void k_geglu(cudaStream_t s, const View& dst, const View& h);
// im2col for the strided / narrow convolutions: col[m][(kh*KW+kw)*C + c], row pitch kpad (zero filled).
void k_im2col(cudaStream_t s, __half* col, int64_t kpad, const View& x, int KW, int KH,
	int s0, int s1, int p0, int p1, int d0, int d1, int64_t OW, int64_t OH);
// conv weight [KW,KH,Cin,Cout] (ggml order) -> [Cout][(kh*KW+kw)*Cin + c], row pitch kpad, f16.
void k_conv_weight_prep(cudaStream_t s, __half* dst, int64_t kpad, const View& w);

// GEGLU projection weights (mlblock_nn.c:159-172, rows [0,D) = value, [D,2D) = gate) re-ordered so that every group of 32
// output columns holds 16 value columns followed by their 16 gate columns: the GEMM epilogue then gates in registers.
// dst/src: 2D rows of `row_elems` elements (f16 weight rows, or row_elems = 1 for the f32 bias).
void k_geglu_rows_prep(cudaStream_t s, void* dst, const View& src, int64_t D);

// ---- attention (attention.cu)
// o[d,q,h,b] = softmax_k(scale * q.k)[.] v ; q:[d,nq,H,B] k:[d,nk,H,B] v given as [nk,d,H,B] (ggml V^T view)
void k_attention(cudaStream_t s, const View& o, const View& q, const View& k, const View& v, float scale, bool causal);

// tcgen05 flash attention (attn_tc.cu): non-causal, d <= 128, f16 token-major operands
struct AttnTC;
bool    attn_tc_supported(const View& o, const View& q, const View& k, const View& v, bool causal);
AttnTC* attn_tc_prepare(const View& o, const View& q, const View& k, const View& v, float scale);
void    attn_tc_launch(cudaStream_t s, AttnTC* a);
void    attn_tc_free(AttnTC* a);
void    attn_tc_set_trace(AttnTC* a, long long* dev_buf);   // debug timeline of CTA 0 (tools/attn_trace.cu)

// ---- tensor-core GEMM / implicit conv (gemm_tc.cu)
struct GemmEpilogue {
	const float* bias = nullptr;       // [N]
	const void*  rowvec = nullptr;     // per-image vector added to every row of that image: [N] x images
	DT           rowvec_dt = DT_F32;
	int64_t      rowvec_stride = 0;    // elements between images
	int64_t      rows_per_image = 0;   // M rows covered by one rowvec
	const void*  residual = nullptr;   // [M,N] same layout as C
	DT           residual_dt = DT_F16;
	int64_t      ldr = 0;
	UnaryOp      act = U_NONE;         // applied after bias/rowvec, before residual
	bool         geglu = false;        // C[M, N/2]: out[:, 16j+i] = h[:, 32j+i] * gelu(h[:, 32j+16+i]) (weights/bias permuted by k_geglu_rows_prep)
	// GroupNorm statistics of the output wanted by its consumer (mlblock_nn.c:78): [images][groups][4] fixed-point words, zeroed
	// before the run. A launch that can accumulate them in its epilogue says so through gemm_tc_gn_fused().
	unsigned long long* gn_stats = nullptr;
	int          gn_groups = 0;
	int64_t      gn_rows_per_image = 0; // rows of C per image (plain GEMM)
};

struct GemmTC;  // opaque prepared launch (tensor maps etc.)
// C[M,N] (row pitch ldc) = A[M,K] (row pitch lda, f16) . B[N,K]^T (row pitch ldb, f16)
GemmTC* gemm_tc_prepare(const __half* A, int64_t lda, const __half* B, int64_t ldb,
	void* C, DT c_dt, int64_t ldc, int64_t M, int64_t N, int64_t K, const GemmEpilogue& ep, int sm_count);
// 3x3 stride-1 pad-1 convolution on channels-last x[N,H,W,Cin] (f16), weights [Cout][9*Cin] prepared.
GemmTC* conv3x3_tc_prepare(const __half* x, int64_t n_img, int64_t H, int64_t W, int64_t Cin,
	const __half* Wt, void* C, DT c_dt, int64_t Cout, const GemmEpilogue& ep, int sm_count);
// direct 3x3 s1 p1 convolution for 3 / 4 output channels (UNet conv_out, last VAE convolution): f16 NHWC in and out, bias only
bool k_conv3x3_small_supported(int64_t Cin, int64_t Cout);
void k_conv3x3_small(cudaStream_t s, __half* y, const __half* x, const __half* w, const float* bias, int64_t N, int64_t H, int64_t W, int64_t Cin, int64_t Cout);
void gemm_tc_launch(cudaStream_t s, GemmTC* g);
bool gemm_tc_gn_fused(const GemmTC* g);   // the launch adds the GroupNorm statistics of its output to ep.gn_stats
void gemm_tc_free(GemmTC* g);
bool gemm_tc_supported(int64_t M, int64_t N, int64_t K);

}  // namespace b200
