// gemm_tc.cu -- the tensor-core contraction kernel of the engine (sm_100a only).
//
// One kernel family serves every linear layer, 1x1 convolution and 3x3 convolution of the
// UNet / VAE / CLIP / TAE graphs (reference emit sites: mlb_nn_linear mlblock_nn.c:16-28,
// mlb_nn_conv2d mlblock_nn.c:31-55):
//     C[M,N] = A[M,K] . B[N,K]^T   (f16 x f16 -> f32 accumulate, the reference's rounding points)
//   * A (activations, rows = tokens/pixels, K contiguous) and B (weights, K contiguous) are staged
//     by TMA (cp.async.bulk.tensor, 128B swizzle) into a multi-stage shared-memory ring;
//   * a single elected thread issues tcgen05.mma (UMMA 128 x BN x 16, cta_group::1), accumulating
//     in tensor memory (TMEM); tcgen05.commit releases ring slots / signals the epilogue;
//   * four epilogue warps read the accumulator with tcgen05.ld and apply the fused epilogue:
//     + bias[N], + per-image vector (the resnet time-embedding add, mlblock_nn.c:139-144),
//     activation, + residual (mlblock_nn.c:154, unet.c:143), store f16/f32.
//   * 3x3 stride-1 pad-1 convolution is an implicit GEMM: the A tile of tap (kh,kw) is a shifted
//     4-D TMA box over the channels-last activation; out-of-bounds rows/cols are zero-filled by
//     the TMA unit, which implements the padding. No im2col buffer exists.
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = epilogue (TMEM lane quarter = warp_id % 4).
#include "kernels.h"
#include "tc_common.cuh"
#include <algorithm>

namespace b200 {

constexpr int BM = 128;
constexpr int BK = 64;            // 64 f16 = 128 B = one swizzle row
constexpr int TMEM_COLS = 256;    // accumulator columns allocated per CTA (>= max BN)
constexpr int MAX_STAGES = 8;
constexpr int EPI_MAX_IMG = 4;    // per-image epilogue vectors staged in shared memory (tile spans <= 4 images)

struct GemmParams {
	int M, N, K, num_kb, BN, stages;
	int conv;                      // 0: plain GEMM, 1: implicit 3x3 s1 p1
	int H, W, Cin, n_img, bw, bh, bi, tiles_w, tiles_h;
	void* C; int c_dt; long long ldc;
	const float* bias;
	const void* rowvec; int rowvec_dt; long long rowvec_stride; long long rows_per_image;
	const void* residual; int residual_dt; long long ldr;
	int act;
	int m_tiles, n_tiles, num_tiles;   // persistent kernel: static tile schedule
	// thread-block cluster of cm x cn CTAs working on cm m-tiles x cn n-tiles: every A tile is loaded once per
	// cluster row (each of its cn CTAs fetches 1/cn of the rows and TMA-multicasts them), every B tile once per column
	int cm, cn, a_rows, b_rows, a_split_dim, a_split_ext, m_ctiles, n_ctiles;
	int two_sm;                        // tcgen05 cta_group::2: a CTA pair computes a 256 x BN tile, each SM loading its 128 rows of A and HALF of B
	int geglu;                         // epilogue gates column pairs: output has N/2 columns (weights pre-permuted)
	int n_stg;                         // staging tiles (2: the TMA store of tile i drains while tile i+1 is written)
	long long* trace;                  // debug timeline of CTA 0's epilogue (GGML_B200_GEMM_TRACE), null in production
	uint32_t mul_nct, mul_tw, mul_th;  // reciprocal multipliers of n_ctiles / tiles_w / tiles_h: tile coordinates without integer division
	int split_k, kbps;                 // split-K (gemm_tc_kernel only): blockIdx.z covers k-blocks [z * kbps, (z + 1) * kbps) and writes f32 partials
	// GroupNorm statistics of the OUTPUT, accumulated by the epilogue of the persistent kernel (mlblock_nn.c:78: the consumer's
	// group_norm then needs no pass of its own over the tensor): [images][groups][4] fixed-point 64-bit words, see gn_fix_add
	unsigned long long* gn_stats; int gn_groups, gn_cpg, gn_ppi;      // gn_ppi: rows of one image inside a tile
	long long rows_per_image_gn;       // plain GEMM: rows of one image (a multiple of the tile height)
	int gn_nimg_total;                 // images in the output
};

// What the reduction pass of a split-K launch needs: the real output and the epilogue that the GEMM kernel skipped.
struct SplitK {
	int S = 1; float* ws = nullptr;
	void* C = nullptr; long long ldc = 0;
	const float* bias = nullptr; const void* rowvec = nullptr; int rowvec_dt = 0; long long rowvec_stride = 0, rows_per_image = 1;
	const void* residual = nullptr; long long ldr = 0; int act = 0;
};

// q = n / d through one multiply-high: mul = ceil(2^32 / d) is exact while n * d < 2^32 (tile counts are far below); d == 1 has mul == 0
static uint32_t fastdiv_mul(uint32_t d) { return d <= 1 ? 0u : (uint32_t)((0x100000000ull + d - 1) / d); }
__device__ __forceinline__ uint32_t fastdiv(uint32_t n, uint32_t mul) { return mul ? __umulhi(n, mul) : n; }

constexpr int GN_TILE_GROUPS = 40;  // groups a 256-column tile can touch (channels per group >= 8)
struct GemmTC {
	CUtensorMap tmA, tmB, tmC, tmR;     // tmC / tmR: output store / residual load maps of the persistent kernel
	GemmParams p;
	dim3 grid;
	size_t smem;
	bool persistent = false;
	SplitK sk;
};

// K-major, 128B-swizzled operand tile: rows of 128 B, 8-row swizzle atoms 1024 B apart.
// (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48),
//  layout SWIZZLE_128B=2 [61,64))
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr)
{
	return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// cute::UMMA::InstrDescriptor: c_format F32=1 [4,6), a/b format F16=0, K-major, N>>3 [17,23), M>>4 [24,29)
__device__ __forceinline__ uint32_t make_idesc(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24); }

__device__ __forceinline__ float act_apply_tc(int op, float x)
{
	switch (op) {
	case U_RELU: return fmaxf(x, 0.0f);
	case U_SILU: return x / (1.0f + __expf(-x));
	case U_GELU: { float u = 0.79788456080286535588f * x * (1.0f + 0.044715f * x * x); return 0.5f * x * (1.0f + tanhf(u)); }
	case U_GELU_QUICK: return x / (1.0f + __expf(-1.702f * x));
	case U_TANH: return tanhf(x);
	default: return x;
	}
}

__device__ __forceinline__ float gelu_tanh_fast(float g)
{
	const float u = 0.79788456080286535588f * g * fmaf(0.044715f * g, g, 1.0f);
	float t; asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
	return 0.5f * g * (1.0f + t);
}

// (x + b) * gelu_tanh(g) with xb_half = b / 2: 7 FMA-pipe operations + one MUFU.TANH per output
//   gelu(g) = g/2 * (1 + tanh(g * (k0 + k1 g^2))),  k0 = sqrt(2/pi), k1 = 0.044715 k0
__device__ __forceinline__ float geglu_gate(float x, float xb_half, float g)
{
	const float t = fmaf(g * g, 0.79788456080286535588f * 0.044715f, 0.79788456080286535588f);
	float th; asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(g * t));
	const float w = fmaf(x, 0.5f, xb_half) * g;
	return fmaf(w, th, w);
}

// fixed-point split of a double (same format as kernels_elem.cu: integer part, fraction * 2^40): integer additions commute,
// so sums accumulated by atomics do not depend on arrival order
// A tile's share of a group sum is a few thousand f16 values: f32 holds it to 1e-7 relative, and the split of an f32 is exact
// (the fraction of an f32 times 2^40 is an integer). No FP64 and no 64-bit float conversions in the epilogue.
__device__ __forceinline__ void gn_fix_split(float v, unsigned long long& hi, unsigned long long& lo)
{
	const float f = floorf(v);
	hi = (unsigned long long)(long long)f; lo = (unsigned long long)(long long)((v - f) * 1099511627776.0f);
}

// ------------------------------------------------------------------ the kernel
__global__ void __launch_bounds__(192, 2)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p)
{
	extern __shared__ __align__(1024) uint8_t smem_raw[];
	uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
	const uint32_t a_bytes = BM * BK * 2, b_bytes = (uint32_t)p.BN * BK * 2, stage_bytes = a_bytes + b_bytes;
	uint64_t* full_bar  = (uint64_t*)(smem + (size_t)p.stages * stage_bytes);
	uint64_t* empty_bar = full_bar + MAX_STAGES;
	uint64_t* accum_bar = empty_bar + MAX_STAGES;
	uint32_t* tmem_slot = (uint32_t*)(accum_bar + 1);
	float* epi_vec = (float*)(tmem_slot + 6);     // [EPI_MAX_IMG][BN] bias + per-image vector, 16-byte aligned

	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int n0 = blockIdx.x * p.BN;
	const int mt = blockIdx.y;
	int m0 = mt * BM, tw0 = 0, th0 = 0, ti0 = 0;
	if (p.conv) {
		int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, ti = mt / (p.tiles_w * p.tiles_h);
		tw0 = tw * p.bw; th0 = th * p.bh; ti0 = ti * p.bi;
	}

	if (threadIdx.x == 0) {
		tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB);
		for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
		mbar_init(accum_bar, 1);
		fence_barrier_init();
	}
	if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem_base = *tmem_slot;

	// split-K: this CTA covers k-blocks [kb0, kb1) and writes its f32 partial sums to slice blockIdx.z of the workspace
	const int kb0 = p.split_k > 1 ? (int)blockIdx.z * p.kbps : 0, kb1 = p.split_k > 1 ? min(p.num_kb, kb0 + p.kbps) : p.num_kb;
	if (warp == 0) {
		// ===== TMA producer =====
		if (lane == 0) {
			const int cpt = p.conv ? p.Cin / BK : 1;  // channel chunks per tap
			for (int kb = kb0; kb < kb1; ++kb) {
				const int s = (kb - kb0) % p.stages;
				const uint32_t ph = (uint32_t)((kb - kb0) / p.stages) & 1;
				mbar_wait(&empty_bar[s], ph ^ 1);
				uint8_t* sa = smem + (size_t)s * stage_bytes;
				uint8_t* sb = sa + a_bytes;
				mbar_expect_tx(&full_bar[s], stage_bytes);
				if (p.conv) {
					const int tap = kb / cpt, cc = kb - tap * cpt;
					const int kh = tap / 3, kw = tap - kh * 3;
					tma_load_4d(sa, &tmA, &full_bar[s], cc * BK, tw0 + kw - 1, th0 + kh - 1, ti0);
				} else {
					tma_load_2d(sa, &tmA, &full_bar[s], kb * BK, m0);
				}
				tma_load_2d(sb, &tmB, &full_bar[s], kb * BK, n0);
			}
		}
	} else if (warp == 1) {
		// ===== MMA issuer (one thread) =====
		if (lane == 0) {
			const uint32_t idesc = make_idesc(p.BN);
			for (int kb = kb0; kb < kb1; ++kb) {
				const int s = (kb - kb0) % p.stages;
				const uint32_t ph = (uint32_t)((kb - kb0) / p.stages) & 1;
				mbar_wait(&full_bar[s], ph);
				tc_fence_after();
				const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
				const uint64_t adesc = make_smem_desc(sa), bdesc = make_smem_desc(sa + a_bytes);
				#pragma unroll
				for (int k = 0; k < BK / 16; ++k) {
					// advance 16 elements (32 B) along K inside the swizzle row: +2 in 16-byte units
					umma_f16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, ((kb - kb0) | k) ? 1u : 0u);
				}
				umma_commit(&empty_bar[s]);          // slot reusable once these MMAs retire
			}
			umma_commit(accum_bar);                  // accumulator complete
		}
	} else {
		// ===== epilogue warps 2..5 =====
		const int quarter = warp & 3;                // TMEM lane quarter this warp may access
		const int r = quarter * 32 + lane;           // row inside the tile
		const int et = threadIdx.x - 64;             // 0..127 among the epilogue threads
		long long grow; bool row_ok; int ii = 0;
		if (p.conv) {
			int wi = r % p.bw, hi = (r / p.bw) % p.bh; ii = r / (p.bw * p.bh);
			int w = tw0 + wi, h = th0 + hi, im = ti0 + ii;
			row_ok = w < p.W && h < p.H && im < p.n_img;
			grow = ((long long)im * p.H + h) * p.W + w;
		} else {
			grow = (long long)m0 + r;
			row_ok = grow < p.M;
		}
		// While the main loop runs, stage bias (+ the per-image vector) of this tile's columns in shared
		// memory: the epilogue then never waits on global loads for them.
		const int n_img_tile = (p.conv && p.rowvec) ? p.bi : 1;
		const bool staged = n_img_tile <= EPI_MAX_IMG;
		if (staged) {
			for (int e = et; e < n_img_tile * p.BN; e += 128) {
				const int im_l = e / p.BN, c = e - im_l * p.BN, col = n0 + c;
				float v = 0.f;
				if (col < p.N) {
					if (p.bias) v = __ldg(p.bias + col);
					if (p.rowvec) {
						const long long im = min((long long)(p.conv ? ti0 + im_l : 0), (long long)p.n_img - 1);
						const long long o = im * p.rowvec_stride + col;
						v += p.rowvec_dt == DT_F16 ? __half2float(((const __half*)p.rowvec)[o]) : ((const float*)p.rowvec)[o];
					}
				}
				epi_vec[e] = v;
			}
			asm volatile("bar.sync 1, 128;" ::: "memory");     // epilogue warps only
		}
		const float* my_vec = epi_vec + ((staged && n_img_tile > 1) ? ii * p.BN : 0);
		const long long img = (!staged && p.rowvec) ? grow / p.rows_per_image : 0;
		const bool has_vec = p.bias || p.rowvec;

		mbar_wait(accum_bar, 0);
		tc_fence_after();
		const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16);
		// residual of the next chunk is fetched while the current one is processed
		const bool res16 = p.residual && p.residual_dt == DT_F16;
		const __half* rrow = res16 ? (const __half*)p.residual + grow * p.ldr + n0 : nullptr;
		const bool res_vec = res16 && row_ok && ((((uintptr_t)rrow) & 15) == 0);
		uint4 rnext[2];
		auto fetch_res = [&](int c0) {
			if (res_vec && n0 + c0 + 16 <= p.N) { rnext[0] = *reinterpret_cast<const uint4*>(rrow + c0); rnext[1] = *reinterpret_cast<const uint4*>(rrow + c0 + 8); }
		};
		fetch_res(0);
		for (int c0 = 0; c0 < p.BN; c0 += 16) {
			uint32_t v[16];
			tmem_ld16(trow + (uint32_t)c0, v);
			uint4 rcur[2] = { rnext[0], rnext[1] };
			if (c0 + 16 < p.BN) fetch_res(c0 + 16);
			tmem_ld_wait();
			const int col0 = n0 + c0;
			if (!row_ok || col0 >= p.N) continue;
			float f[16];
			#pragma unroll
			for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
			const bool full = col0 + 16 <= p.N;
			if (has_vec) {
				if (staged) {
					#pragma unroll
					for (int j = 0; j < 16; j += 4) { float4 t = *reinterpret_cast<const float4*>(my_vec + c0 + j); f[j] += t.x; f[j+1] += t.y; f[j+2] += t.z; f[j+3] += t.w; }
				} else {
					#pragma unroll
					for (int j = 0; j < 16; ++j) if (full || col0 + j < p.N) {
						if (p.bias) f[j] += __ldg(p.bias + col0 + j);
						if (p.rowvec) { long long o = img * p.rowvec_stride + col0 + j;
							f[j] += p.rowvec_dt == DT_F16 ? __half2float(((const __half*)p.rowvec)[o]) : ((const float*)p.rowvec)[o]; }
					}
				}
			}
			if (p.act != U_NONE) {
				#pragma unroll
				for (int j = 0; j < 16; ++j) f[j] = act_apply_tc(p.act, f[j]);
			}
			if (p.residual) {
				const long long ro = grow * p.ldr + col0;
				if (res16) {
					if (res_vec && full) {
						const __half2* ha = reinterpret_cast<const __half2*>(&rcur[0]); const __half2* hb = reinterpret_cast<const __half2*>(&rcur[1]);
						#pragma unroll
						for (int j = 0; j < 4; ++j) { float2 x = __half22float2(ha[j]), y = __half22float2(hb[j]);
							f[2*j] += x.x; f[2*j+1] += x.y; f[8+2*j] += y.x; f[8+2*j+1] += y.y; }
					} else {
						const __half* rp = (const __half*)p.residual + ro;
						#pragma unroll
						for (int j = 0; j < 16; ++j) if (full || col0 + j < p.N) f[j] += __half2float(rp[j]);
					}
				} else {
					const float* rp = (const float*)p.residual + ro;
					#pragma unroll
					for (int j = 0; j < 16; ++j) if (full || col0 + j < p.N) f[j] += rp[j];
				}
			}
			const long long co = grow * p.ldc + col0 + (p.split_k > 1 ? (long long)blockIdx.z * p.M * p.ldc : 0);
			if (p.c_dt == DT_F16) {
				__half* cp = (__half*)p.C + co;
				if (full && ((co & 7) == 0)) {
					uint4 a, b; __half2* ha = reinterpret_cast<__half2*>(&a); __half2* hb = reinterpret_cast<__half2*>(&b);
					#pragma unroll
					for (int j = 0; j < 4; ++j) { ha[j] = __floats2half2_rn(f[2*j], f[2*j+1]); hb[j] = __floats2half2_rn(f[8+2*j], f[8+2*j+1]); }
					*reinterpret_cast<uint4*>(cp) = a; *reinterpret_cast<uint4*>(cp + 8) = b;
				} else {
					#pragma unroll
					for (int j = 0; j < 16; ++j) if (full || col0 + j < p.N) cp[j] = __float2half_rn(f[j]);
				}
			} else {
				float* cp = (float*)p.C + co;
				if (full && ((co & 3) == 0)) {
					#pragma unroll
					for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(cp + 4 * j) = make_float4(f[4*j], f[4*j+1], f[4*j+2], f[4*j+3]);
				} else {
					#pragma unroll
					for (int j = 0; j < 16; ++j) if (full || col0 + j < p.N) cp[j] = f[j];
				}
			}
		}
	}
	tc_fence_before();
	__syncthreads();
	if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, TMEM_COLS); }
}

// ------------------------------------------------------------------ the persistent kernel (v2)
// One CTA per SM walks a static list of output tiles (n fastest, so CTAs of a wave share activation rows in L2):
//   warps 0..7  epilogue (TMEM lane quarter = warp % 4, column half = warp / 4): accumulator -> registers (tcgen05.ld) -> + bias / per-image
//               vector (staged in shared memory), activation, + residual (TMA-loaded into the staging tile) ->
//               f16 -> 64B-swizzled staging tile -> TMA store. Global traffic of the epilogue is bulk-async only.
//   warp 8      TMA producer, warp 9 MMA issuer: both run warp-uniform control flow with one elected lane issuing,
//               so descriptors live in uniform registers (no R2UR per instruction); highest warp ids = issue priority.
// The accumulator is double-buffered in tensor memory (2 x 256 columns): the epilogue of tile i overlaps the main
// loop of tile i+1, and barrier setup / TMEM allocation / descriptor prefetch happen once per SM, not per tile.
constexpr int P_THREADS = 320;                           // 8 epilogue warps + TMA producer + MMA issuer
constexpr int P_EPI_MAX_IMG = 8;
constexpr int STG_CHUNK_COLS = 32;                        // staging chunk: [128 rows][32 f16] = 8 KB, SWIZZLE_64B
constexpr int STG_CHUNK_BYTES = BM * STG_CHUNK_COLS * 2;

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, const void* src, int c0, int c1)
{
	asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
		:: "l"(tm), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* tm, const void* src, int c0, int c1, int c2, int c3)
{
	asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
		:: "l"(tm), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, uint16_t mask)
{
	asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
		:: "r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
__device__ __forceinline__ void tma_load_4d_mc(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2, int c3, uint16_t mask)
{
	asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
		:: "r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "h"(mask) : "memory");
}
// arrive on the same barrier offset in every CTA of `mask` once all previously issued MMAs have completed
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask)
{
	asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
		:: "r"(smem_u32(bar)), "h"(mask) : "memory");
}
// ---- cta_group::2 (CTA pair) forms. TMA loads issued by either CTA complete on the LEADER's barrier (same offset,
// CTA-rank bits of the shared::cluster address cleared); MMA / commit are issued by the leader for both SMs.
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1)
{
	asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
		:: "r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2, int c3)
{
	asm volatile("cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
		:: "r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
	asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
		"tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
		:: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar, uint16_t mask)
{
	asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
		:: "r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* dst_smem, uint32_t cols)
{
	asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(dst_smem)), "r"(cols) : "memory");
	asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t addr, uint32_t cols)
{ asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(addr), "r"(cols) : "memory"); }
// arrive on the barrier at the same offset in CTA `rank` of the cluster. No .release.cluster qualifier: that form costs a
// GPU-scope memory barrier (MEMBAR.ALL.GPU + ERRBAR) per arrival, and what the arrival orders here are tensor-memory
// reads that tcgen05.wait::ld + tcgen05.fence::before_thread_sync have already completed.
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank)
{
	asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\t"
		"mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" :: "r"(smem_u32(bar)), "r"(rank) : "memory");
}
__device__ __forceinline__ uint32_t make_idesc_m(int m, int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }

__device__ __forceinline__ void cluster_sync_all()
{ asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_id_x() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_count_x() { uint32_t r; asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r)); return r; }

// TWO_SM: tcgen05 cta_group::2 CTA pair (needs a cluster launch); ACT: 0 none, 1 SiLU, 2 run-time p.act, 3 GEGLU gate; HAS_RES: f16
// residual tile added in the epilogue; GNS: GroupNorm statistics of the output (p.gn_stats) -- compile time, so that the launches
// without a group_norm behind them keep the lean epilogue (the run-time form cost them 30 registers and 1-8 %)
template <int ACT, bool HAS_RES, bool TWO_SM, bool GNS = false>
__global__ void __launch_bounds__(P_THREADS, 1)
gemm_tc_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
	const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR, const GemmParams p)
{
	extern __shared__ __align__(1024) uint8_t smem_raw[];
	uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
	const uint32_t a_bytes = BM * BK * 2, b_bytes = (uint32_t)(TWO_SM ? p.BN / 2 : p.BN) * BK * 2, stage_bytes = a_bytes + b_bytes;
	constexpr int ACC_PER_CHUNK = ACT == 3 ? 2 * STG_CHUNK_COLS : STG_CHUNK_COLS;   // accumulator columns behind one staging chunk
	const int n_chunks = (p.BN + ACC_PER_CHUNK - 1) / ACC_PER_CHUNK;
	uint8_t* stg = smem + (size_t)p.stages * stage_bytes;                       // staging tile (1024-aligned)
	float* epi_vec = (float*)(stg + (size_t)p.n_stg * n_chunks * STG_CHUNK_BYTES);   // [P_EPI_MAX_IMG][BN]
	uint64_t* full_bar  = (uint64_t*)(epi_vec + P_EPI_MAX_IMG * 256);
	uint64_t* empty_bar = full_bar + MAX_STAGES;
	uint64_t* acc_full  = empty_bar + MAX_STAGES;      // [2]
	uint64_t* acc_empty = acc_full + 2;                // [2]
	uint64_t* res_full  = acc_empty + 2;               // [1]
	uint32_t* tmem_slot = (uint32_t*)(res_full + 1);
	float2* gn_part = (float2*)(tmem_slot + 4);        // GNS: [images of a tile][256 epilogue threads] (sum, sum of squares) of a column pair

	const int warp = warp_id_uniform(), lane = threadIdx.x & 31;
	if (p.trace && blockIdx.x == 0 && threadIdx.x == 0) p.trace[6] = clock64();          // kernel entry
	const int csize = TWO_SM ? 2 : p.cm * p.cn;
	const int crank = csize > 1 ? (int)cluster_ctarank() : 0;
	const int pr = TWO_SM ? crank : 0;               // rank inside the CTA pair (0 = leader, issues the MMAs)
	const int rn = TWO_SM ? 0 : crank % p.cn, rm = TWO_SM ? 0 : crank / p.cn;
	const int cluster = csize > 1 ? (int)cluster_id_x() : (int)blockIdx.x, nclusters = csize > 1 ? (int)cluster_count_x() : (int)gridDim.x;
	const int num_ctiles = p.m_ctiles * p.n_ctiles;
	// CTAs sharing my A tile (same cluster row) / my B tile (same cluster column)
	uint16_t row_mask = 0, col_mask = 0;
	for (int j = 0; j < p.cn; ++j) row_mask |= (uint16_t)(1u << (rm * p.cn + j));
	for (int i = 0; i < p.cm; ++i) col_mask |= (uint16_t)(1u << (i * p.cn + rn));
	if (threadIdx.x == 0) {
		tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); tma_prefetch_desc(&tmC);
		if (p.residual) tma_prefetch_desc(&tmR);
		// a ring slot is free once every CTA that multicasts into it has seen ALL its readers release it
		for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], p.cm + p.cn - 1); }
		// CTA pair: the leader's accumulator-empty barrier collects the epilogue warps of BOTH CTAs
		for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], TWO_SM ? 16 : 8); }
		mbar_init(res_full, 1);
		fence_barrier_init();
	}
	if (warp == 9) { if (TWO_SM) tmem_alloc_2sm(tmem_slot, 512); else tmem_alloc(tmem_slot, 512); }
	tc_fence_before();
	__syncthreads();
	if (csize > 1) cluster_sync_all();                 // peers' barriers exist before anything is multicast to them
	tc_fence_after();
	const uint32_t tmem_base = uniform_u32(*tmem_slot);

	// cluster tile id -> this CTA's tile coordinates (n fastest). Tiles past the edge (odd tile counts) are processed
	// like any other: their loads are zero-filled and their stores clipped by the TMA unit.
	auto tile_coords = [&](int tile, int& n0, int& m0, int& tw0, int& th0, int& ti0) {
		const int mq = (int)fastdiv((uint32_t)tile, p.mul_nct);                  // tile / n_ctiles
		const int nt = (tile - mq * p.n_ctiles) * p.cn + rn, mt = mq * (TWO_SM ? 2 : p.cm) + rm + pr;
		n0 = nt * p.BN; m0 = mt * BM; tw0 = th0 = ti0 = 0;
		if (p.conv) {
			const int q1 = (int)fastdiv((uint32_t)mt, p.mul_tw), ti = (int)fastdiv((uint32_t)q1, p.mul_th);   // mt / tiles_w, / tiles_h
			tw0 = (mt - q1 * p.tiles_w) * p.bw; th0 = (q1 - ti * p.tiles_h) * p.bh; ti0 = ti * p.bi;
		}
	};

	if (warp == 8) {
		// ===== TMA producer =====
		const int cpt = p.conv ? p.Cin / BK : 1;          // channel chunks per filter tap
		int s = 0; uint32_t ph = 0;                        // ring slot and its phase, carried across tiles (no division in the loop)
		for (int tile = cluster; tile < num_ctiles; tile += nclusters) {
			int n0, m0, tw0, th0, ti0; tile_coords(tile, n0, m0, tw0, th0, ti0);
			int tap = 0, cc = 0;
			for (int kb = 0; kb < p.num_kb; ++kb) {
				mbar_wait(&empty_bar[s], ph ^ 1);
				if (TWO_SM) {
					// both CTAs of the pair load their own A rows and their half of the B rows; every byte is accounted
					// on the leader's barrier, which is the one the MMA issuer waits on
					if (elect_one()) {
						uint8_t* sa = smem + (size_t)s * stage_bytes;
						if (pr == 0) mbar_expect_tx(&full_bar[s], 2 * stage_bytes);
						if (p.conv) {
							const int kh = (tap * 11) >> 5, kw = tap - kh * 3;
							tma_load_4d_2sm(sa, &tmA, &full_bar[s], cc * BK, tw0 + kw - 1, th0 + kh - 1, ti0);
						} else tma_load_2d_2sm(sa, &tmA, &full_bar[s], kb * BK, m0);
						tma_load_2d_2sm(sa + a_bytes, &tmB, &full_bar[s], kb * BK, n0 + pr * (p.BN / 2));
					}
				} else
				if (elect_one()) {
					uint8_t* sa = smem + (size_t)s * stage_bytes;
					mbar_expect_tx(&full_bar[s], stage_bytes);      // my slice + the slices the peers multicast to me
					uint8_t* sa_slice = sa + rn * p.a_rows * 128;
					uint8_t* sb_slice = sa + a_bytes + rm * p.b_rows * 128;
					if (p.conv) {
						const int kh = (tap * 11) >> 5, kw = tap - kh * 3;      // tap / 3 for tap < 9
						int cw = tw0 + kw - 1, chh = th0 + kh - 1, ci = ti0;
						if (p.a_split_dim == 3) ci += rn * p.a_split_ext; else if (p.a_split_dim == 2) chh += rn * p.a_split_ext; else cw += rn * p.a_split_ext;
						if (p.cn > 1) tma_load_4d_mc(sa_slice, &tmA, &full_bar[s], cc * BK, cw, chh, ci, row_mask);
						else tma_load_4d(sa_slice, &tmA, &full_bar[s], cc * BK, cw, chh, ci);
					} else {
						if (p.cn > 1) tma_load_2d_mc(sa_slice, &tmA, &full_bar[s], kb * BK, m0 + rn * p.a_rows, row_mask);
						else tma_load_2d(sa_slice, &tmA, &full_bar[s], kb * BK, m0);
					}
					if (p.cm > 1) tma_load_2d_mc(sb_slice, &tmB, &full_bar[s], kb * BK, n0 + rm * p.b_rows, col_mask);
					else tma_load_2d(sb_slice, &tmB, &full_bar[s], kb * BK, n0);
				}
				__syncwarp();
				if (++cc == cpt) { cc = 0; ++tap; }
				if (++s == p.stages) { s = 0; ph ^= 1; }
			}
		}
	} else if (warp == 9) {
		// ===== MMA issuer =====
		const uint32_t idesc = TWO_SM ? make_idesc_m(2 * BM, p.BN) : make_idesc(p.BN);
		const uint64_t adesc0 = make_smem_desc(smem_u32(smem)), bdesc0 = make_smem_desc(smem_u32(smem) + a_bytes);
		const uint32_t stage16 = stage_bytes >> 4;
		int s = 0; uint32_t ph = 0, lt = 0;                // ring slot / phase, local tile counter
		if (!(TWO_SM && pr != 0))                        // CTA pair: only the leader issues (for both tensor cores)
		for (int tile = cluster; tile < num_ctiles; tile += nclusters, ++lt) {
			const uint32_t buf = lt & 1;
			mbar_wait(&acc_empty[buf], ((lt >> 1) & 1) ^ 1);       // epilogue drained this accumulator
			tc_fence_after();
			const uint32_t td = tmem_base + buf * 256;
			for (int kb = 0; kb < p.num_kb; ++kb) {
				mbar_wait(&full_bar[s], ph);
				tc_fence_after();
				if (elect_one()) {
					const uint64_t ad = adesc0 + (uint64_t)(s * stage16), bd = bdesc0 + (uint64_t)(s * stage16);
					if (TWO_SM) {
						#pragma unroll
						for (int k = 0; k < BK / 16; ++k)
							umma_f16_2sm(td, ad + 2 * k, bd + 2 * k, idesc, (kb | k) ? 1u : 0u);
						umma_commit_2sm(&empty_bar[s], 3);           // the slot is free in both CTAs
						if (kb == p.num_kb - 1) umma_commit_2sm(&acc_full[buf], 3);
					} else {
						#pragma unroll
						for (int k = 0; k < BK / 16; ++k)
							umma_f16(td, ad + 2 * k, bd + 2 * k, idesc, (kb | k) ? 1u : 0u);
						if (csize > 1) umma_commit_mc(&empty_bar[s], (uint16_t)(row_mask | col_mask));
						else umma_commit(&empty_bar[s]);
						if (kb == p.num_kb - 1) umma_commit(&acc_full[buf]);
					}
				}
				__syncwarp();
				if (++s == p.stages) { s = 0; ph ^= 1; }
			}
		}
	} else {
		// ===== epilogue warps 0..7: TMEM lane quarter = warp % 4, column half = warp / 4 =====
		const int quarter = warp & 3, grp = warp >> 2;
		const int r = quarter * 32 + lane;                 // row of the tile = TMEM lane
		const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
		const uint32_t row_sw = (uint32_t)((r >> 1) & 3);  // SWIZZLE_64B: 16-byte unit index ^= address bits [7,8]
		const uint32_t stg_bytes = (uint32_t)n_chunks * STG_CHUNK_BYTES;
		const uint32_t my_row0 = smem_u32(stg) + r * 64;   // shared-space address of this row in chunk 0 of staging tile 0
		const bool has_vec = p.bias || p.rowvec;
		const int n_img_tile = (p.conv && p.rowvec) ? p.bi : 1;
		const int ii = p.conv ? r / (p.bw * p.bh) : 0;     // image of this row inside the tile
		const int half_chunks = (n_chunks + 1) >> 1;
		const int c_lo = grp * half_chunks * ACC_PER_CHUNK, c_hi = min(p.BN, (grp + 1) * half_chunks * ACC_PER_CHUNK);
		const int n_out = ACT == 3 ? p.N / 2 : p.N;         // output columns
		auto out_col0 = [&](int n0) { return (ACT == 3 ? n0 / 2 : n0) + warp * STG_CHUNK_COLS; };   // first output column of this warp's chunk
		const int et = threadIdx.x;                        // 0..255 among the epilogue threads
		const bool chunk_owner = lane == 0 && warp < n_chunks;    // lane 0 of warp w stores (and re-fills) staging chunk w
		uint8_t* my_chunk0 = stg + warp * STG_CHUNK_BYTES;
		auto bias_of = [&](int e, int n0, int ti0) {       // bias + per-image vector of staged element e
			const int im_l = n_img_tile > 1 ? e / p.BN : 0, c = e - im_l * p.BN, col = n0 + c;
			float v = 0.f;
			if (col < p.N) {
				if (p.bias) v = __ldg(p.bias + col);
				if (p.rowvec) {
					const long long im = min((long long)(p.conv ? ti0 + im_l : 0), (long long)p.n_img - 1);
					const long long o = im * p.rowvec_stride + col;
					v += p.rowvec_dt == DT_F16 ? __half2float(((const __half*)p.rowvec)[o]) : ((const float*)p.rowvec)[o];
				}
			}
			if (ACT == 3 && !(c & 16)) v *= 0.5f;          // GEGLU value columns: the gate's factor 1/2 is folded into value and bias
			return v;
		};
		auto res_load = [&](int tile, uint32_t sbuf) {     // chunk owners: residual chunk of `tile` -> staging tile sbuf
			int n0, m0, tw0, th0, ti0; tile_coords(tile, n0, m0, tw0, th0, ti0);
			uint8_t* dst = my_chunk0 + sbuf * stg_bytes;
			if (p.conv) tma_load_4d(dst, &tmR, res_full, n0 + warp * STG_CHUNK_COLS, tw0, th0, ti0);
			else tma_load_2d(dst, &tmR, res_full, n0 + warp * STG_CHUNK_COLS, m0);
		};
		float bias_next = 0.f;                             // first staged element of the next tile, fetched a tile ahead
		if (cluster < num_ctiles) {
			int n0, m0, tw0, th0, ti0; tile_coords(cluster, n0, m0, tw0, th0, ti0);
			if (has_vec && et < n_img_tile * p.BN) bias_next = bias_of(et, n0, ti0);
			if (HAS_RES) {
				if (et == 0) mbar_expect_tx(res_full, (uint32_t)n_chunks * STG_CHUNK_BYTES);
				if (chunk_owner) res_load(cluster, 0);
			}
		}
		// GroupNorm statistics of the output (GNS): after a tile is staged, warp w sums the values and squares of staging chunk w,
		// one column pair per lane (lane & 15) over the rows of one parity (lane >> 4), per image of the tile, and leaves them in
		// shared memory. Behind the next tile's first barrier one thread per (image, group) adds up the pairs of its group in a
		// fixed order and sends the sums to global memory as fixed-point integer atomics: bit-reproducible, and no
		// shared-memory atomics (64-bit ones are CAS loops: 30 us per launch when 40 lanes meet on one word).
		int gn_n0 = 0, gn_img0 = 0; bool gn_pending = false;
		const int gn_nimg = GNS ? BM / p.gn_ppi : 0;
		auto gn_flush = [&]() {
			if (!gn_pending) return;
			const int g_first = gn_n0 / p.gn_cpg, c_end = min(p.BN, p.N - gn_n0);
			for (int e = et; e < gn_nimg * GN_TILE_GROUPS; e += 256) {
				const int im = e / GN_TILE_GROUPS, gl = e - im * GN_TILE_GROUPS, g = g_first + gl, img = gn_img0 + im;
				const int c_lo = max(g * p.gn_cpg - gn_n0, 0), c_hi = min((g + 1) * p.gn_cpg - gn_n0, c_end);
				if (g >= p.gn_groups || img >= p.gn_nimg_total || c_lo >= c_hi) continue;
				float a = 0.f, b = 0.f;
				for (int c = c_lo; c < c_hi; c += 2) {
					const int t0 = (c >> 5) * 32 + ((c & 31) >> 1);            // thread that summed this pair's even rows; +16: odd rows
					const float2 x = gn_part[im * 256 + t0], y = gn_part[im * 256 + t0 + 16];
					a += x.x + y.x; b += x.y + y.y;
				}
				unsigned long long* dst = p.gn_stats + ((size_t)img * p.gn_groups + g) * 4;
				unsigned long long h, l;
				gn_fix_split(a, h, l); atomicAdd(dst, h); atomicAdd(dst + 1, l);
				gn_fix_split(b, h, l); atomicAdd(dst + 2, h); atomicAdd(dst + 3, l);
			}
			gn_pending = false;
		};
		uint32_t lt = 0;
		for (int tile = cluster; tile < num_ctiles; tile += nclusters, ++lt) {
			int n0, m0, tw0, th0, ti0; tile_coords(tile, n0, m0, tw0, th0, ti0);
			const uint32_t buf = lt & 1, sbuf = lt & (uint32_t)(p.n_stg - 1);
			const uint32_t my_row = my_row0 + sbuf * stg_bytes;
			uint8_t* my_chunk = my_chunk0 + sbuf * stg_bytes;
			const int next = tile + nclusters;
			#define GEMM_TR(ev) do { if (p.trace && blockIdx.x == 0 && threadIdx.x == 0 && lt < 16) p.trace[lt * 8 + (ev)] = clock64(); } while (0)
			GEMM_TR(0);
			// stage bias (+ per-image vector) of this tile's columns
			if (has_vec) {
				if (et < n_img_tile * p.BN) epi_vec[et] = bias_next;
				for (int e = et + 256; e < n_img_tile * p.BN; e += 256) epi_vec[e] = bias_of(e, n0, ti0);
			}
			// the previous tile's stores must have left the staging tile (with a residual the owners waited before re-filling it)
			if (!HAS_RES && chunk_owner) { if (p.n_stg == 2) tma_store_wait_read1(); else tma_store_wait_read(); }
			asm volatile("bar.sync 1, 256;" ::: "memory");
			if (GNS) gn_flush();                           // the previous tile's sums: complete behind this barrier
			if (has_vec && next < num_ctiles && et < n_img_tile * p.BN) {
				int n1, m1, tw1, th1, ti1; tile_coords(next, n1, m1, tw1, th1, ti1);
				bias_next = bias_of(et, n1, ti1);          // in flight during the tile
			}
			const uint32_t my_vec = smem_u32(epi_vec) + (n_img_tile > 1 ? ii * p.BN : 0) * 4;
			GEMM_TR(1);
			mbar_wait(&acc_full[buf], (lt >> 1) & 1);
			if (HAS_RES) mbar_wait(res_full, lt & 1);
			tc_fence_after();
			GEMM_TR(2);
			const uint32_t trow = tmem_base + buf * 256 + lane_off;
			auto process = [&](const uint32_t* v, int c0) {
				float f[16];
				#pragma unroll
				for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
				if (has_vec) {
					#pragma unroll
					for (int j = 0; j < 16; j += 4) { const float4 t = lds128f(my_vec + (c0 + j) * 4); f[j] += t.x; f[j+1] += t.y; f[j+2] += t.z; f[j+3] += t.w; }
				}
				if (ACT == 1) {
					#pragma unroll
					for (int j = 0; j < 16; ++j) f[j] = __fdividef(f[j], 1.0f + __expf(-f[j]));
				} else if (ACT == 2) {
					#pragma unroll
					for (int j = 0; j < 16; ++j) f[j] = act_apply_tc(p.act, f[j]);
				}
				const uint32_t chunk = my_row + (uint32_t)(c0 / STG_CHUNK_COLS) * STG_CHUNK_BYTES;
				const uint32_t u0 = (uint32_t)((c0 % STG_CHUNK_COLS) >> 3);     // 0 or 2
				const uint32_t s0 = chunk + ((u0 ^ row_sw) << 4), s1 = chunk + (((u0 + 1) ^ row_sw) << 4);
				if (HAS_RES) {
					const uint4 ra = lds128(s0), rb = lds128(s1);
					const __half2* ha = reinterpret_cast<const __half2*>(&ra); const __half2* hb = reinterpret_cast<const __half2*>(&rb);
					#pragma unroll
					for (int j = 0; j < 4; ++j) { float2 x = __half22float2(ha[j]), y = __half22float2(hb[j]);
						f[2*j] += x.x; f[2*j+1] += x.y; f[8+2*j] += y.x; f[8+2*j+1] += y.y; }
				}
				uint4 a, b; __half2* pa = reinterpret_cast<__half2*>(&a); __half2* pb = reinterpret_cast<__half2*>(&b);
				#pragma unroll
				for (int j = 0; j < 4; ++j) { pa[j] = __floats2half2_rn(f[2*j], f[2*j+1]); pb[j] = __floats2half2_rn(f[8+2*j], f[8+2*j+1]); }
				sts128(s0, a); sts128(s1, b);
			};
			// GEGLU: 16 value columns and their 16 gate columns -> 16 gated outputs
			auto process_geglu = [&](const uint32_t* vx, const uint32_t* vg, int c0) {
				float f[16];
				#pragma unroll
				for (int j = 0; j < 16; j += 4) {
					float4 bx = make_float4(0.f, 0.f, 0.f, 0.f), bg = bx;        // bx holds HALF the value bias (bias_of)
					if (has_vec) { bx = lds128f(my_vec + (c0 + j) * 4); bg = lds128f(my_vec + (c0 + 16 + j) * 4); }
					f[j]     = geglu_gate(__uint_as_float(vx[j]),     bx.x, __uint_as_float(vg[j])     + bg.x);
					f[j + 1] = geglu_gate(__uint_as_float(vx[j + 1]), bx.y, __uint_as_float(vg[j + 1]) + bg.y);
					f[j + 2] = geglu_gate(__uint_as_float(vx[j + 2]), bx.z, __uint_as_float(vg[j + 2]) + bg.z);
					f[j + 3] = geglu_gate(__uint_as_float(vx[j + 3]), bx.w, __uint_as_float(vg[j + 3]) + bg.w);
				}
				const int oc = c0 >> 1;                            // output column inside the tile
				const uint32_t chunk = my_row + (uint32_t)(oc / STG_CHUNK_COLS) * STG_CHUNK_BYTES;
				const uint32_t u0 = (uint32_t)((oc % STG_CHUNK_COLS) >> 3);
				uint4 a, b; __half2* pa = reinterpret_cast<__half2*>(&a); __half2* pb = reinterpret_cast<__half2*>(&b);
				#pragma unroll
				for (int j = 0; j < 4; ++j) { pa[j] = __floats2half2_rn(f[2*j], f[2*j+1]); pb[j] = __floats2half2_rn(f[8+2*j], f[8+2*j+1]); }
				sts128(chunk + ((u0 ^ row_sw) << 4), a); sts128(chunk + (((u0 + 1) ^ row_sw) << 4), b);
			};
			if (ACT == 3) {
				if (c_lo < c_hi) {
					uint32_t va[16], vb[16], vc[16], vd[16];
					tmem_ld16(trow + (uint32_t)c_lo, va); tmem_ld16(trow + (uint32_t)(c_lo + 16), vb);
					for (int c0 = c_lo; c0 < c_hi; c0 += 64) {
						tmem_ld_wait();
						if (c0 + 32 < c_hi) { tmem_ld16(trow + (uint32_t)(c0 + 32), vc); tmem_ld16(trow + (uint32_t)(c0 + 48), vd); }
						process_geglu(va, vb, c0);
						if (c0 + 32 < c_hi) {
							tmem_ld_wait();
							if (c0 + 64 < c_hi) { tmem_ld16(trow + (uint32_t)(c0 + 64), va); tmem_ld16(trow + (uint32_t)(c0 + 80), vb); }
							process_geglu(vc, vd, c0 + 32);
						}
					}
				}
			} else
			// 16 accumulator columns at a time, the next tcgen05.ld in flight while the current ones are processed
			if (c_lo < c_hi) {
				uint32_t va[16], vb[16];
				tmem_ld16(trow + (uint32_t)c_lo, va);
				for (int c0 = c_lo; c0 < c_hi; c0 += 32) {
					tmem_ld_wait();
					if (c0 + 16 < c_hi) tmem_ld16(trow + (uint32_t)(c0 + 16), vb);
					process(va, c0);
					if (c0 + 16 < c_hi) {
						tmem_ld_wait();
						if (c0 + 32 < c_hi) tmem_ld16(trow + (uint32_t)(c0 + 32), va);
						process(vb, c0 + 16);
					}
				}
			}
			GEMM_TR(3);
			// accumulator drained: hand the TMEM buffer back to the MMA warp (one arrival per warp)
			tc_fence_before();
			__syncwarp();
			if (lane == 0) { if (TWO_SM && pr != 0) mbar_arrive_remote(&acc_empty[buf], 0); else mbar_arrive(&acc_empty[buf]); }
			fence_proxy_async();                           // staging writes -> visible to the TMA (async proxy)
			asm volatile("bar.sync 1, 256;" ::: "memory");
			GEMM_TR(4);
			if (HAS_RES && et == 0 && next < num_ctiles) mbar_expect_tx(res_full, (uint32_t)n_chunks * STG_CHUNK_BYTES);
			if (chunk_owner) {
				if (out_col0(n0) < n_out) {
					if (p.conv) tma_store_4d(&tmC, my_chunk, out_col0(n0), tw0, th0, ti0);
					else tma_store_2d(&tmC, my_chunk, out_col0(n0), m0);
				}
				tma_store_commit();
				// with one staging tile the residual of the next tile lands where the statistics pass below still reads: load it afterwards
				if (HAS_RES && next < num_ctiles && !(GNS && p.n_stg == 1)) { if (p.n_stg == 2) tma_store_wait_read1(); else tma_store_wait_read(); res_load(next, (lt + 1) & (uint32_t)(p.n_stg - 1)); }
			}
			GEMM_TR(5);
			if (GNS) {
				const int img0 = p.conv ? ti0 : (int)(m0 / p.rows_per_image_gn);
				const bool tile_ok = p.conv ? ti0 < p.gn_nimg_total : m0 < p.M;
				const int col = warp * STG_CHUNK_COLS + (lane & 15) * 2;          // column pair inside the tile
				if (tile_ok && warp < n_chunks && n0 + col < p.N) {
					const uint32_t cbase = smem_u32(my_chunk), u = (uint32_t)(lane & 15) >> 2, w4 = (uint32_t)(lane & 3) * 4;
					for (int im = 0; im < gn_nimg; ++im) {
						float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
						#pragma unroll 8
						for (int rr = im * p.gn_ppi + (lane >> 4); rr < (im + 1) * p.gn_ppi; rr += 2) {
							uint32_t hv;
							asm volatile("ld.shared.b32 %0, [%1];" : "=r"(hv) : "r"(cbase + (uint32_t)rr * 64 + ((u ^ ((uint32_t)(rr >> 1) & 3)) << 4) + w4));
							const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hv));
							s0 += f.x; q0 = fmaf(f.x, f.x, q0); s1 += f.y; q1 = fmaf(f.y, f.y, q1);
						}
						gn_part[im * 256 + et] = make_float2(s0 + s1, q0 + q1);
					}
				}
				gn_n0 = n0; gn_img0 = img0; gn_pending = tile_ok;
				if (HAS_RES && p.n_stg == 1) {
					__syncwarp();                              // the whole warp has read its chunk
					if (chunk_owner && next < num_ctiles) { tma_store_wait_read(); res_load(next, 0); }
				}
			}
		}
		if (GNS) { asm volatile("bar.sync 1, 256;" ::: "memory"); gn_flush(); }
		if (chunk_owner) tma_store_wait_all();
	}
	tc_fence_before();
	__syncthreads();
	if (csize > 1) cluster_sync_all();                 // no CTA leaves while a peer may still signal its barriers
	if (p.trace && blockIdx.x == 0 && threadIdx.x == 0) p.trace[7] = clock64();          // kernel exit
	if (warp == 9) { tc_fence_after(); if (TWO_SM) tmem_dealloc_2sm(tmem_base, 512); else tmem_dealloc(tmem_base, 512); }
}

// ------------------------------------------------------------------ split-K reduction + epilogue
// out[row, col] = act(sum_s ws[s][row][col] + bias[col] + rowvec[image(row)][col]) + residual[row, col]   (f16, 4 columns per thread)
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ ws, int S, long long M, int N, __half* __restrict__ C, long long ldc,
	const float* __restrict__ bias, const void* __restrict__ rowvec, int rowvec_dt, long long rowvec_stride, long long rows_per_image,
	int act, const __half* __restrict__ residual, long long ldr)
{
	const int n4 = N >> 2;
	const long long total = M * n4, slice = M * (long long)N;
	for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
		const long long row = idx / n4; const int col = (int)(idx - row * n4) * 4;
		const float* src = ws + row * N + col;
		float4 a = *reinterpret_cast<const float4*>(src);
		for (int sidx = 1; sidx < S; ++sidx) { const float4 b = *reinterpret_cast<const float4*>(src + sidx * slice); a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
		float f[4] = { a.x, a.y, a.z, a.w };
		if (bias) { const float4 b = *reinterpret_cast<const float4*>(bias + col); f[0] += b.x; f[1] += b.y; f[2] += b.z; f[3] += b.w; }
		if (rowvec) {
			const long long o = (row / rows_per_image) * rowvec_stride + col;
			#pragma unroll
			for (int j = 0; j < 4; ++j) f[j] += rowvec_dt == DT_F16 ? __half2float(((const __half*)rowvec)[o + j]) : ((const float*)rowvec)[o + j];
		}
		if (act != U_NONE) {
			#pragma unroll
			for (int j = 0; j < 4; ++j) f[j] = act_apply_tc(act, f[j]);
		}
		if (residual) {
			#pragma unroll
			for (int j = 0; j < 4; ++j) f[j] += __half2float(residual[row * ldr + col + j]);
		}
		__half2 h0 = __floats2half2_rn(f[0], f[1]), h1 = __floats2half2_rn(f[2], f[3]);
		__half* dst = C + row * ldc + col;
		*reinterpret_cast<__half2*>(dst) = h0; *reinterpret_cast<__half2*>(dst + 2) = h1;
	}
}

// ------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
	const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode()
{
	static PFN_encodeTiled fn = nullptr;
	if (!fn) {
		void* p = nullptr; cudaDriverEntryPointQueryResult qres;
		CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
		if (!p || qres != cudaDriverEntryPointSuccess) B200_FATAL("cuTensorMapEncodeTiled not available in this driver");
		fn = (PFN_encodeTiled)p;
	}
	return fn;
}

static void encode_map(CUtensorMap* tm, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
	const cuuint32_t* box, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B)
{
	cuuint32_t es[5] = {1, 1, 1, 1, 1};
	CUresult r = get_encode()(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes,
		box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
		CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if (r != CUDA_SUCCESS) {
		B200_FATAL("cuTensorMapEncodeTiled failed (%d): rank %d dims %llu %llu %llu %llu box %u %u %u %u base %p stride0 %llu", (int)r, rank,
			(unsigned long long)dims[0], (unsigned long long)dims[1], rank > 2 ? (unsigned long long)dims[2] : 0ull, rank > 3 ? (unsigned long long)dims[3] : 0ull,
			box[0], box[1], rank > 2 ? box[2] : 0u, rank > 3 ? box[3] : 0u, base, (unsigned long long)strides_bytes[0]);
	}
}

bool gemm_tc_supported(int64_t M, int64_t N, int64_t K)
{
	return M >= 1 && N >= 1 && K >= 8 && (K % 8) == 0;
}

// Pick the N tile: minimise (waves x per-tile time), per-tile time ~ BN + fixed overhead.
static int pick_bn(int64_t m_tiles, int64_t N, int sm_count)
{
	int best = 64; double best_cost = 1e30;
	for (int bn = 256; bn >= 16; bn -= 16) {
		if (bn > 16 && bn - 16 >= N) continue;   // no point in a tile much wider than N
		int64_t nt = (N + bn - 1) / bn, tiles = nt * m_tiles;
		int64_t waves = (tiles + sm_count - 1) / sm_count;
		double cost = (double)waves * (bn + 40.0);
		if (cost < best_cost - 1e-9) { best_cost = cost; best = bn; }
	}
	return best;
}

static void finish_setup(GemmTC* g, const GemmEpilogue& ep, int64_t m_tiles, int sm_count)
{
	GemmParams& p = g->p;
	// Two CTAs per SM (2 x 256 TMEM columns): while one CTA drains its accumulator through the epilogue
	// the other keeps the tensor pipe and the TMA queue busy. Shared memory is sized so that exactly
	// two fit (>= 77 KB each, so never three: a third could not allocate tensor memory).
	p.BN = pick_bn(m_tiles, p.N, sm_count * 2);
	size_t stage = (size_t)BM * BK * 2 + (size_t)p.BN * BK * 2;
	int stages = (int)std::min<size_t>(MAX_STAGES, (108 * 1024) / stage);
	stages = std::max(2, std::min(stages, std::max(2, p.num_kb)));
	p.stages = stages;
	g->smem = std::max<size_t>(stages * stage + 1024 /*align*/ + (2 * MAX_STAGES + 1) * 8 + 32 + EPI_MAX_IMG * 256 * 4, 77 * 1024);
	g->grid = dim3((unsigned)((p.N + p.BN - 1) / p.BN), (unsigned)m_tiles);
	p.bias = ep.bias; p.rowvec = ep.rowvec; p.rowvec_dt = ep.rowvec_dt; p.rowvec_stride = ep.rowvec_stride;
	p.rows_per_image = ep.rows_per_image > 0 ? ep.rows_per_image : 1;
	p.residual = ep.residual; p.residual_dt = ep.residual_dt; p.ldr = ep.ldr; p.act = ep.act;
}

// ---- persistent kernel setup
static bool env_on(const char* n, bool dflt) { const char* e = getenv(n); return e && *e ? atoi(e) != 0 : dflt; }

// Split-K for contractions with few output tiles and a long K (the 8x8-level convolutions of SD1.x: M = 1024, N = 1280,
// K = 11520 / 23040): with 40-112 tiles every CTA has to pull 5+ MB of operands through its own L2 port and the launch is
// ingest-bound at ~590 TFLOP/s. Here 128 x 256 tiles are cut along K so that about one CTA per SM works on 1/S of it,
// writing f32 partials; a second pass sums them and applies the epilogue.
static bool persistent_eligible(const GemmParams& p, const GemmEpilogue& ep);
static int max_active_clusters(int csize, int sm_count);
// The same few-tile, long-K contractions fit ONE wave of CTA pairs (tcgen05 cta_group::2, 256 x BN per pair) when BN is chosen so that
// about 3/4 of the pairs get a tile: every SM then pulls its 128 rows of A and only half of a NARROW B tile per k-block, no f32
// partials and no reduction pass. Measured (run r3j, M = 1024, N = 1280): K = 11520 37.1 us with BN = 96 against 46.0 us for split-K
// (38.9 us with BN = 128), K = 23040 67.6 us against 71.8 us. Returns the BN to use, 0 if there is no such tiling.
static int one_wave_pair_bn(const GemmParams& p, const GemmEpilogue& ep, int64_t m_tiles, int sm_count)
{
	if (!env_on("GGML_B200_GEMM_2SM", true) || !env_on("GGML_B200_GEMM_PAIRWAVE", true) || m_tiles < 2 || !persistent_eligible(p, ep) || ep.geglu) return 0;
	const int nclusters = max_active_clusters(2, sm_count);
	if (nclusters <= 0) return 0;
	const int64_t m_pairs = (m_tiles + 1) / 2;
	for (int bn = 96; bn <= 128; bn += 32) {
		const int64_t n_tiles = (p.N + bn - 1) / bn, ctiles = m_pairs * n_tiles;
		if (n_tiles >= 2 && ctiles <= nclusters && ctiles * 4 >= (int64_t)nclusters * 2) return bn;
	}
	return 0;
}
static int choose_split_k(const GemmParams& p, const GemmEpilogue& ep, int64_t m_tiles, int sm_count)
{
	if (!env_on("GGML_B200_GEMM_SPLITK", true)) return 1;
	if (ep.geglu || p.c_dt != DT_F16 || (p.N % 4) || (p.ldc % 4) || ((uintptr_t)p.C & 7)) return 1;
	if (ep.residual && (ep.residual_dt != DT_F16)) return 1;
	if (ep.bias && ((uintptr_t)ep.bias & 15)) return 1;
	const int64_t tiles = m_tiles * ((p.N + 255) / 256);
	if (tiles * 2 > sm_count || p.num_kb < 128) return 1;       // measured: K = 11520 -7 %, K = 23040 -18 %, but K = 5120 +25 % (the partial sums cost more than they save)
	if (one_wave_pair_bn(p, ep, m_tiles, sm_count)) return 1;    // one wave of CTA pairs beats the split (run r3j)
	int S = (int)std::min<int64_t>(8, sm_count / tiles);
	S = std::min(S, p.num_kb / 16);
	return S >= 2 ? S : 1;
}

static void finish_setup_split_k(GemmTC* g, const GemmEpilogue& ep, int64_t m_tiles, int S)
{
	GemmParams& p = g->p;
	SplitK& k = g->sk;
	k.S = S; k.C = p.C; k.ldc = p.ldc;
	k.bias = ep.bias; k.rowvec = ep.rowvec; k.rowvec_dt = ep.rowvec_dt; k.rowvec_stride = ep.rowvec_stride;
	k.rows_per_image = ep.rows_per_image > 0 ? ep.rows_per_image : 1;
	k.residual = ep.residual; k.ldr = ep.ldr; k.act = ep.act;
	CUDA_CHECK(cudaMalloc(&k.ws, (size_t)S * p.M * p.N * sizeof(float)));
	p.split_k = S; p.kbps = (p.num_kb + S - 1) / S;
	p.BN = (int)std::min<int64_t>(256, (p.N + 15) / 16 * 16);
	const size_t stage = (size_t)BM * BK * 2 + (size_t)p.BN * BK * 2;
	p.stages = (int)std::max<size_t>(2, std::min<size_t>(MAX_STAGES, (196 * 1024) / stage));
	g->smem = p.stages * stage + 1024 + (2 * MAX_STAGES + 1) * 8 + 32 + EPI_MAX_IMG * 256 * 4;
	g->grid = dim3((unsigned)((p.N + p.BN - 1) / p.BN), (unsigned)m_tiles, (unsigned)S);
	// the GEMM kernel writes raw f32 partial sums; everything else happens in the reduction pass
	p.C = k.ws; p.c_dt = DT_F32; p.ldc = p.N;
	p.bias = nullptr; p.rowvec = nullptr; p.residual = nullptr; p.act = U_NONE; p.rows_per_image = 1;
	if (getenv("GGML_B200_GEMM_DEBUG"))
		B200_LOG("gemm M=%d N=%d K=%d conv=%d m_tiles=%lld -> split-K %d x %d k-blocks, BN=%d, grid %u x %u x %u, stages %d", p.M, p.N, p.K, p.conv,
			(long long)m_tiles, S, p.kbps, p.BN, g->grid.x, g->grid.y, g->grid.z, p.stages);
}

static bool persistent_eligible(const GemmParams& p, const GemmEpilogue& ep)
{
	if (!env_on("GGML_B200_GEMM_PERSISTENT", true)) return false;
	if (p.c_dt != DT_F16 || (p.ldc % 8) || ((uintptr_t)p.C & 15)) return false;
	if (ep.residual && (ep.residual_dt != DT_F16 || ep.ldr != p.ldc || ((uintptr_t)ep.residual & 15))) return false;
	if (p.conv && ep.rowvec && p.bi > P_EPI_MAX_IMG) return false;
	if (ep.geglu && (p.conv || ep.residual || ep.rowvec || ep.act != U_NONE || (p.N % 64))) return false;
	return true;
}

// How many clusters of `csize` CTAs (1 CTA per SM, ~200 KB shared memory each) the device can hold at once.
static int max_active_clusters(int csize, int sm_count);

// Tile / cluster choice of the persistent kernel: minimise waves x per-tile cycles. Per tile
//   tensor pipe : num_kb * 4 MMAs of max(BN/2 [tcgen05 floor], 32 + BN/4 [A+B shared-memory reads at 128 B/clk]) cycles
//   L2 -> SM    : (128 + BN) * 128 B per k-block and CTA at ~58 B/clk per SM when every SM pulls (measured ~17 TB/s
//                 chip-wide): a lone 128 x 256 tile is ingest-bound at ~60 % of the tensor peak. TMA multicast in a
//                 cm x cn cluster only trims this (every SM still receives whole tiles); the CTA pair of
//                 tcgen05 cta_group::2 halves the B share per SM and is what lifts the big contractions to >80 %
//   epilogue    : ~5 cycles per column, overlaps the next tile unless it is the longest of the three.
struct TileChoice { int bn, cm, cn, two_sm; };
static TileChoice pick_tiles_persistent(int64_t m_tiles, int64_t N, int num_kb, int sm_count, bool geglu)
{
	static const int cands[][2] = { {1, 1}, {2, 1}, {1, 2}, {2, 2}, {4, 1}, {4, 2} };
	const char* e = getenv("GGML_B200_GEMM_CLUSTER");
	const int max_cluster = e && *e ? atoi(e) : 8;
	const bool allow_2sm = env_on("GGML_B200_GEMM_2SM", true);
	TileChoice best = {0, 1, 1, 0}; double best_cost = 1e30;
	auto bn_ok = [&](int bn, int64_t n_tiles) {
		if (bn - 16 >= N) return false;
		if (n_tiles > 1 && (bn % (geglu ? 2 * STG_CHUNK_COLS : STG_CHUNK_COLS))) return false;   // staging chunks must not straddle tiles
		if (geglu && (bn % 32)) return false;
		return true;
	};
	for (auto& c : cands) {
		const int cm = c[0], cn = c[1], cs = cm * cn;
		if (cs > max_cluster || cm > m_tiles) continue;
		const int nclusters = cs == 1 ? sm_count : max_active_clusters(cs, sm_count);
		if (nclusters <= 0) continue;
		for (int bn = 256; bn >= 16; bn -= 16) {
			const int64_t n_tiles = (N + bn - 1) / bn;
			if (!bn_ok(bn, n_tiles)) continue;
			if (cn > n_tiles || (bn % (8 * cm))) continue;            // B slices are whole 8-row swizzle atoms
			const int64_t ctiles = ((m_tiles + cm - 1) / cm) * ((n_tiles + cn - 1) / cn), waves = (ctiles + nclusters - 1) / nclusters;
			const double mma = (double)num_kb * 4 * std::max(bn / 2.0, 32.0 + bn / 4.0), epi = 250.0 + 5.0 * bn;
			// measured: multicast relieves the L2 slices but every SM still ingests the full tiles -- about a quarter of
			// the shared operand's cost goes away, not (cluster-1)/cluster of it
			const double l2 = (double)num_kb * (128.0 * (cn > 1 ? 0.75 : 1.0) + (double)bn * (cm > 1 ? 0.75 : 1.0)) * 128.0 / 58.0;
			const double cost = ((double)waves * std::max(std::max(mma, l2), epi) + epi) * (1.0 + 0.01 * (cs - 1));   // ties -> smaller cluster
			if (cost < best_cost - 1e-9) { best_cost = cost; best = {bn, cm, cn, 0}; }
		}
	}
	// CTA pair (tcgen05 cta_group::2): 256 x BN per pair at the same BN/2 cycles per MMA, each SM pulling its A rows and
	// only half of B: the way to stay under the per-SM L2 bandwidth with wide tiles
	if (allow_2sm && m_tiles >= 2 && max_cluster >= 2) {
		const int nclusters = max_active_clusters(2, sm_count);
		for (int bn = 256; nclusters > 0 && bn >= 32; bn -= 16) {
			const int64_t n_tiles = (N + bn - 1) / bn;
			if (!bn_ok(bn, n_tiles)) continue;
			const int64_t ctiles = ((m_tiles + 1) / 2) * n_tiles, waves = (ctiles + nclusters - 1) / nclusters;
			const double mma = (double)num_kb * 4 * std::max(bn / 2.0, 32.0 + bn / 8.0), epi = 250.0 + 5.0 * bn;
			const double l2 = (double)num_kb * (128.0 + bn / 2.0) * 128.0 / 58.0;
			const double cost = ((double)waves * std::max(std::max(mma, l2), epi) + epi) * 1.01;
			if (cost < best_cost - 1e-9) { best_cost = cost; best = {bn, 1, 1, 1}; }
		}
	}
	return best;
}

static void finish_setup_persistent(GemmTC* g, const GemmEpilogue& ep, int64_t m_tiles, int sm_count)
{
	GemmParams& p = g->p;
	p.geglu = ep.geglu ? 1 : 0;
	TileChoice tc = pick_tiles_persistent(m_tiles, p.N, p.num_kb, sm_count, ep.geglu);
	{	// few tiles, long K: the one-wave CTA-pair tiling (the cost model above underrates it: it ranks by waves x ingest only)
		const int64_t tiles256 = m_tiles * ((p.N + 255) / 256);
		if (tiles256 * 2 <= sm_count && p.num_kb >= 128) { const int bn = one_wave_pair_bn(p, ep, m_tiles, sm_count); if (bn) tc = {bn, 1, 1, 1}; }
	}
	if (const char* f = getenv("GGML_B200_GEMM_FORCE")) {      // "bn,cm,cn[,two_sm]": tuning experiments (tools/gemm_bench.py)
		int bn = 0, cm = 1, cn = 1, two = 0;
		if (sscanf(f, "%d,%d,%d,%d", &bn, &cm, &cn, &two) >= 3 && bn >= 16 && bn <= 256 && bn % 16 == 0 && bn % (8 * cm) == 0 &&
			(bn % STG_CHUNK_COLS == 0 || bn >= p.N) && cm * cn <= 8 && max_active_clusters(cm * cn, sm_count) > 0) tc = {bn, two ? 1 : cm, two ? 1 : cn, two ? 1 : 0};
	}
	if (getenv("GGML_B200_GEMM_TRACE")) { CUDA_CHECK(cudaMalloc(&p.trace, 16 * 8 * 8)); CUDA_CHECK(cudaMemset(p.trace, 0, 16 * 8 * 8)); }
	if (getenv("GGML_B200_GEMM_DEBUG"))
		B200_LOG("gemm M=%d N=%d K=%d conv=%d m_tiles=%lld -> BN=%d cluster %dx%d%s", p.M, p.N, p.K, p.conv, (long long)m_tiles, tc.bn, tc.cm, tc.cn, tc.two_sm ? " cta_group::2" : "");
	p.BN = tc.bn; p.cm = tc.cm; p.cn = tc.cn; p.two_sm = tc.two_sm;
	p.m_tiles = (int)m_tiles; p.n_tiles = (p.N + p.BN - 1) / p.BN; p.num_tiles = p.m_tiles * p.n_tiles;
	p.m_ctiles = (p.m_tiles + p.cm - 1) / p.cm; p.n_ctiles = (p.n_tiles + p.cn - 1) / p.cn;
	p.a_rows = BM / p.cn; p.b_rows = p.BN / p.cm;
	if (p.two_sm) { p.m_ctiles = (p.m_tiles + 1) / 2; p.b_rows = p.BN / 2; }
	p.mul_nct = fastdiv_mul((uint32_t)p.n_ctiles); p.mul_tw = fastdiv_mul((uint32_t)std::max(1, p.tiles_w)); p.mul_th = fastdiv_mul((uint32_t)std::max(1, p.tiles_h));
	const size_t stage = (size_t)BM * BK * 2 + (size_t)(p.two_sm ? p.BN / 2 : p.BN) * BK * 2;
	const size_t acc_per_chunk = ep.geglu ? 2 * STG_CHUNK_COLS : STG_CHUNK_COLS;
	const size_t n_chunks = (p.BN + acc_per_chunk - 1) / acc_per_chunk;
	// A second staging tile lets the TMA store of tile i drain while tile i+1 is written, but the operand ring needs
	// ~150 KB in flight to cover the ~2500-cycle loaded L2 latency at 57 B/clk: only take it when 5 stages still fit.
	p.n_stg = 1;
	// GroupNorm statistics of the output in the epilogue: whole tiles inside the tensor (no clipped rows), rows of an image
	// contiguous inside a tile and even in number, channels per group even (a column pair never straddles two groups)
	p.gn_stats = nullptr; p.gn_groups = 0; p.gn_cpg = 0; p.gn_ppi = BM; p.rows_per_image_gn = 1; p.gn_nimg_total = 0;
	if (ep.gn_stats && !ep.geglu && ep.gn_groups > 0 && p.N % ep.gn_groups == 0 && env_on("GGML_B200_GN_EPILOGUE", true)) {
		const int cpg = p.N / ep.gn_groups;
		bool ok = cpg >= 8 && cpg % 2 == 0;
		if (p.conv) ok = ok && p.W % p.bw == 0 && p.H % p.bh == 0 && (p.bw * p.bh) % 2 == 0 && p.bi <= P_EPI_MAX_IMG;
		else ok = ok && ep.gn_rows_per_image > 0 && ep.gn_rows_per_image % BM == 0 && p.M % ep.gn_rows_per_image == 0;
		if (ok) {
			p.gn_stats = ep.gn_stats; p.gn_groups = ep.gn_groups; p.gn_cpg = cpg;
			p.gn_ppi = p.conv ? p.bw * p.bh : BM;
			p.rows_per_image_gn = p.conv ? 1 : ep.gn_rows_per_image;
			p.gn_nimg_total = p.conv ? p.n_img : (int)(p.M / ep.gn_rows_per_image);
		}
	}
	const size_t gn_bytes = p.gn_stats ? (size_t)(BM / p.gn_ppi) * 256 * sizeof(float2) : 0;
	auto fixed_bytes = [&](int n_stg) { return 1024 + n_stg * n_chunks * STG_CHUNK_BYTES + (size_t)P_EPI_MAX_IMG * 256 * 4 + (2 * MAX_STAGES + 5) * 8 + 64 + gn_bytes; };
	// GEGLU tiles are epilogue-bound when K is short (the gate costs ~10 cycles per column): there the second staging tile
	// is worth a ring stage
	const int min_stages_2stg = (ep.geglu && p.num_kb <= 8) ? 4 : 5;
	if ((227 * 1024 - fixed_bytes(2)) / stage >= (size_t)min_stages_2stg) p.n_stg = 2;
	if (const char* e = getenv("GGML_B200_GEMM_NSTG")) { const int v = atoi(e); if ((v == 1 || v == 2) && (227 * 1024 - fixed_bytes(v)) / stage >= 2) p.n_stg = v; }
	const size_t fixed = fixed_bytes(p.n_stg);
	int stages = (int)std::min<size_t>(MAX_STAGES, (227 * 1024 - fixed) / stage);
	p.stages = std::max(2, stages);
	g->smem = fixed + p.stages * stage;
	const int cs = p.two_sm ? 2 : p.cm * p.cn;
	const int nclusters = (int)std::min<int64_t>((int64_t)p.m_ctiles * p.n_ctiles, cs == 1 ? sm_count : max_active_clusters(cs, sm_count));
	g->grid = dim3((unsigned)(nclusters * cs));
	if (getenv("GGML_B200_GEMM_DEBUG")) B200_LOG("  grid %u cluster %d smem %zu stages %d n_stg %d max_active_clusters(%d)=%d", g->grid.x, cs, g->smem, p.stages, p.n_stg, cs, cs > 1 ? max_active_clusters(cs, sm_count) : 0);
	g->persistent = true;
	p.bias = ep.bias; p.rowvec = ep.rowvec; p.rowvec_dt = ep.rowvec_dt; p.rowvec_stride = ep.rowvec_stride;
	p.rows_per_image = ep.rows_per_image > 0 ? ep.rows_per_image : 1;
	p.residual = ep.residual; p.residual_dt = ep.residual_dt; p.ldr = ep.ldr; p.act = ep.act;
}

GemmTC* gemm_tc_prepare(const __half* A, int64_t lda, const __half* B, int64_t ldb,
	void* C, DT c_dt, int64_t ldc, int64_t M, int64_t N, int64_t K, const GemmEpilogue& ep, int sm_count)
{
	if (!gemm_tc_supported(M, N, K) || (lda % 8) || (ldb % 8) || ((uintptr_t)A & 15) || ((uintptr_t)B & 15))
		B200_FATAL("gemm_tc_prepare: unsupported operands M=%lld N=%lld K=%lld lda=%lld ldb=%lld", (long long)M, (long long)N, (long long)K, (long long)lda, (long long)ldb);
	GemmTC* g = new GemmTC();
	GemmParams& p = g->p;
	memset(&p, 0, sizeof(p));
	p.M = (int)M; p.N = (int)N; p.K = (int)K; p.num_kb = (int)((K + BK - 1) / BK);
	p.C = C; p.c_dt = c_dt; p.ldc = ldc;
	int64_t m_tiles = (M + BM - 1) / BM;
	const int split = choose_split_k(p, ep, m_tiles, sm_count);
	if (split > 1) finish_setup_split_k(g, ep, m_tiles, split);
	else if (persistent_eligible(p, ep)) {
		finish_setup_persistent(g, ep, m_tiles, sm_count);
		cuuint64_t dc[2] = { (cuuint64_t)(ep.geglu ? N / 2 : N), (cuuint64_t)M }, sc[1] = { (cuuint64_t)ldc * 2 };
		cuuint32_t bc[2] = { STG_CHUNK_COLS, BM };
		encode_map(&g->tmC, C, 2, dc, sc, bc, CU_TENSOR_MAP_SWIZZLE_64B);
		if (ep.residual) encode_map(&g->tmR, ep.residual, 2, dc, sc, bc, CU_TENSOR_MAP_SWIZZLE_64B);
	} else {
		if (ep.geglu) B200_FATAL("gemm_tc_prepare: GEGLU epilogue needs the persistent kernel (M=%lld N=%lld K=%lld)", (long long)M, (long long)N, (long long)K);
		finish_setup(g, ep, m_tiles, sm_count);
	}
	cuuint64_t da[2] = { (cuuint64_t)K, (cuuint64_t)M }, sa[1] = { (cuuint64_t)lda * 2 };
	cuuint32_t ba[2] = { BK, (cuuint32_t)(g->persistent ? p.a_rows : BM) };
	encode_map(&g->tmA, A, 2, da, sa, ba);
	cuuint64_t db[2] = { (cuuint64_t)K, (cuuint64_t)N }, sb[1] = { (cuuint64_t)ldb * 2 };
	cuuint32_t bb[2] = { BK, (cuuint32_t)(g->persistent ? p.b_rows : p.BN) };
	encode_map(&g->tmB, B, 2, db, sb, bb);
	return g;
}

static int pow2_le(int64_t x, int cap) { int r = 1; while (r * 2 <= x && r * 2 <= cap) r *= 2; return r; }
// Tile extent along one image axis: the largest power of two that divides n (no wasted rows);
// if that is tiny compared to n, accept a partial last tile instead.
static int tile_extent(int64_t n, int cap)
{
	int d = 1; while (d * 2 <= cap && n % (d * 2) == 0) d *= 2;
	int le = pow2_le(n, cap);
	return (d >= 8 || d == le) ? d : le;
}

GemmTC* conv3x3_tc_prepare(const __half* x, int64_t n_img, int64_t H, int64_t W, int64_t Cin,
	const __half* Wt, void* C, DT c_dt, int64_t Cout, const GemmEpilogue& ep, int sm_count)
{
	if (Cin % BK || ((uintptr_t)x & 15) || ((uintptr_t)Wt & 15))
		B200_FATAL("conv3x3_tc_prepare: Cin=%lld must be a multiple of %d", (long long)Cin, BK);
	GemmTC* g = new GemmTC();
	GemmParams& p = g->p;
	memset(&p, 0, sizeof(p));
	p.conv = 1; p.H = (int)H; p.W = (int)W; p.Cin = (int)Cin; p.n_img = (int)n_img;
	p.M = (int)(n_img * H * W); p.N = (int)Cout; p.K = (int)(9 * Cin); p.num_kb = (int)(9 * Cin / BK);
	p.C = C; p.c_dt = c_dt; p.ldc = Cout;
	// 128 output pixels per tile = bw x bh x bi (columns x rows x images), powers of two
	p.bw = tile_extent(W, BM);
	p.bh = tile_extent(H, BM / p.bw);
	p.bi = BM / (p.bw * p.bh);
	p.tiles_w = (int)((W + p.bw - 1) / p.bw);
	p.tiles_h = (int)((H + p.bh - 1) / p.bh);
	int64_t tiles_i = (n_img + p.bi - 1) / p.bi;
	int64_t m_tiles = (int64_t)p.tiles_w * p.tiles_h * tiles_i;
	const int split = choose_split_k(p, ep, m_tiles, sm_count);
	if (split > 1) finish_setup_split_k(g, ep, m_tiles, split);
	else if (persistent_eligible(p, ep)) {
		finish_setup_persistent(g, ep, m_tiles, sm_count);
		cuuint64_t dc[4] = { (cuuint64_t)Cout, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n_img };
		cuuint64_t sc[3] = { (cuuint64_t)Cout * 2, (cuuint64_t)W * Cout * 2, (cuuint64_t)H * W * Cout * 2 };
		cuuint32_t bc[4] = { STG_CHUNK_COLS, (cuuint32_t)p.bw, (cuuint32_t)p.bh, (cuuint32_t)p.bi };
		encode_map(&g->tmC, C, 4, dc, sc, bc, CU_TENSOR_MAP_SWIZZLE_64B);
		if (ep.residual) encode_map(&g->tmR, ep.residual, 4, dc, sc, bc, CU_TENSOR_MAP_SWIZZLE_64B);
	} else finish_setup(g, ep, m_tiles, sm_count);
	cuuint64_t da[4] = { (cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n_img };
	cuuint64_t sa[3] = { (cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2 };
	cuuint32_t ba[4] = { BK, (cuuint32_t)p.bw, (cuuint32_t)p.bh, (cuuint32_t)p.bi };
	if (g->persistent && p.cn > 1) {
		// each CTA of a cluster row fetches 1/cn of the tile's pixels: split the slowest box dim that is wide enough
		// (rows of the tile are ordered image, row, column, so such a slice is a contiguous block of tile rows)
		if (p.bi >= p.cn) { p.a_split_dim = 3; p.a_split_ext = p.bi / p.cn; ba[3] = p.a_split_ext; }
		else if (p.bh >= p.cn) { p.a_split_dim = 2; p.a_split_ext = p.bh / p.cn; ba[2] = p.a_split_ext; }
		else { p.a_split_dim = 1; p.a_split_ext = p.bw / p.cn; ba[1] = p.a_split_ext; }
	}
	encode_map(&g->tmA, x, 4, da, sa, ba);
	cuuint64_t db[2] = { (cuuint64_t)(9 * Cin), (cuuint64_t)Cout }, sb[1] = { (cuuint64_t)(9 * Cin) * 2 };
	cuuint32_t bb[2] = { BK, (cuuint32_t)(g->persistent ? p.b_rows : p.BN) };
	encode_map(&g->tmB, Wt, 2, db, sb, bb);
	return g;
}

typedef void (*PersistentKernel)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const GemmParams);
template <bool TWO_SM> static PersistentKernel persistent_variant_t(const GemmParams& p)
{
	const int act = p.act == U_NONE ? 0 : p.act == U_SILU ? 1 : 2;
	const bool res = p.residual != nullptr;
	if (p.geglu) return gemm_tc_persistent_kernel<3, false, TWO_SM>;
	if (p.gn_stats) switch (act * 2 + (res ? 1 : 0)) {
	case 0: return gemm_tc_persistent_kernel<0, false, TWO_SM, true>;
	case 1: return gemm_tc_persistent_kernel<0, true, TWO_SM, true>;
	case 2: return gemm_tc_persistent_kernel<1, false, TWO_SM, true>;
	case 3: return gemm_tc_persistent_kernel<1, true, TWO_SM, true>;
	case 4: return gemm_tc_persistent_kernel<2, false, TWO_SM, true>;
	default: return gemm_tc_persistent_kernel<2, true, TWO_SM, true>;
	}
	switch (act * 2 + (res ? 1 : 0)) {
	case 0: return gemm_tc_persistent_kernel<0, false, TWO_SM>;
	case 1: return gemm_tc_persistent_kernel<0, true, TWO_SM>;
	case 2: return gemm_tc_persistent_kernel<1, false, TWO_SM>;
	case 3: return gemm_tc_persistent_kernel<1, true, TWO_SM>;
	case 4: return gemm_tc_persistent_kernel<2, false, TWO_SM>;
	default: return gemm_tc_persistent_kernel<2, true, TWO_SM>;
	}
}
static PersistentKernel persistent_variant(const GemmParams& p) { return p.two_sm ? persistent_variant_t<true>(p) : persistent_variant_t<false>(p); }
static void persistent_attrs_once()
{
	static bool done = false;
	if (done) return;
	done = true;
	PersistentKernel ks[] = {
		gemm_tc_persistent_kernel<0, false, false>, gemm_tc_persistent_kernel<0, true, false>, gemm_tc_persistent_kernel<1, false, false>,
		gemm_tc_persistent_kernel<1, true, false>, gemm_tc_persistent_kernel<2, false, false>, gemm_tc_persistent_kernel<2, true, false>,
		gemm_tc_persistent_kernel<3, false, false>,
		gemm_tc_persistent_kernel<0, false, true>, gemm_tc_persistent_kernel<0, true, true>, gemm_tc_persistent_kernel<1, false, true>,
		gemm_tc_persistent_kernel<1, true, true>, gemm_tc_persistent_kernel<2, false, true>, gemm_tc_persistent_kernel<2, true, true>,
		gemm_tc_persistent_kernel<3, false, true>,
		gemm_tc_persistent_kernel<0, false, false, true>, gemm_tc_persistent_kernel<0, true, false, true>, gemm_tc_persistent_kernel<1, false, false, true>,
		gemm_tc_persistent_kernel<1, true, false, true>, gemm_tc_persistent_kernel<2, false, false, true>, gemm_tc_persistent_kernel<2, true, false, true>,
		gemm_tc_persistent_kernel<0, false, true, true>, gemm_tc_persistent_kernel<0, true, true, true>, gemm_tc_persistent_kernel<1, false, true, true>,
		gemm_tc_persistent_kernel<1, true, true, true>, gemm_tc_persistent_kernel<2, false, true, true>, gemm_tc_persistent_kernel<2, true, true, true> };
	for (PersistentKernel k : ks) CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
}

void gemm_tc_launch(cudaStream_t s, GemmTC* g)
{
	static bool attr_set = false;
	if (!attr_set) {
		CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
		attr_set = true;
	}
	if (g->persistent) {
		persistent_attrs_once();
		const int cs = g->p.two_sm ? 2 : g->p.cm * g->p.cn;
		cudaLaunchConfig_t cfg = {};
		cfg.gridDim = g->grid; cfg.blockDim = dim3(P_THREADS); cfg.dynamicSmemBytes = g->smem; cfg.stream = s;
		cudaLaunchAttribute at[1];
		at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
		cfg.attrs = at; cfg.numAttrs = cs > 1 ? 1 : 0;
		CUDA_CHECK(cudaLaunchKernelEx(&cfg, persistent_variant(g->p), g->tmA, g->tmB, g->tmC, g->tmR, g->p));
	} else gemm_tc_kernel<<<g->grid, 192, g->smem, s>>>(g->tmA, g->tmB, g->p);
	g_stats.kernel_launches++;
	if (g->sk.S > 1) {
		const SplitK& k = g->sk;
		const long long total = (long long)g->p.M * (g->p.N / 4);
		const unsigned blocks = (unsigned)std::min<long long>((total + 255) / 256, 148LL * 8);
		splitk_reduce_kernel<<<blocks, 256, 0, s>>>(k.ws, k.S, g->p.M, g->p.N, (__half*)k.C, k.ldc, k.bias, k.rowvec, k.rowvec_dt, k.rowvec_stride,
			k.rows_per_image, k.act, (const __half*)k.residual, k.ldr);
		g_stats.kernel_launches++;
	}
}

static int max_active_clusters(int csize, int sm_count)
{
	static int cache[17] = {0};
	if (csize < 1 || csize > 16) return 0;
	if (cache[csize]) return cache[csize];
	persistent_attrs_once();
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3((unsigned)(sm_count / csize * csize)); cfg.blockDim = dim3(P_THREADS); cfg.dynamicSmemBytes = 200 * 1024;
	cudaLaunchAttribute at[1];
	at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
	cfg.attrs = at; cfg.numAttrs = 1;
	int n = 0;
	cudaError_t e = cudaOccupancyMaxActiveClusters(&n, gemm_tc_persistent_kernel<0, false, false>, &cfg);
	if (e != cudaSuccess) { cudaGetLastError(); n = 0; }
	n = std::min(n, sm_count / csize);
	cache[csize] = n > 0 ? n : -1;
	return cache[csize];
}

bool gemm_tc_gn_fused(const GemmTC* g) { return g && g->persistent && g->p.gn_stats != nullptr; }

void gemm_tc_free(GemmTC* g)
{
	if (g->p.trace) {
		long long h[16 * 8];
		cudaDeviceSynchronize();
		cudaMemcpy(h, g->p.trace, sizeof(h), cudaMemcpyDeviceToHost);
		fprintf(stderr, "[ggml_b200] epilogue timeline of CTA 0 (cycles since its first tile; M=%d N=%d K=%d BN=%d): start | bias+store-drain+bar | acc ready | drained | fence+bar | stores issued\n", g->p.M, g->p.N, g->p.K, g->p.BN);
		fprintf(stderr, "  kernel entry %lld, exit %lld (same origin)\n", h[6] - h[0], h[7] - h[0]);
		for (int t = 0; t < 16 && h[t * 8]; ++t) {
			fprintf(stderr, "  tile %2d:", t);
			for (int e = 0; e < 6; ++e) fprintf(stderr, " %8lld", h[t * 8 + e] - h[0]);
			fprintf(stderr, "\n");
		}
		cudaFree(g->p.trace);
	}
	if (g->sk.ws) cudaFree(g->sk.ws);
	delete g;
}

}  // namespace b200
