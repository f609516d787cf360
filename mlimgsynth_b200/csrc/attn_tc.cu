// attn_tc.cu -- flash-style fused attention on the 5th-gen tensor cores (sm_100a).
// Replaces the reference's unfused ggml_nn_attention chain (ggml_extend.c:200-221: QK^T, scale,
// softmax, PV with the [nk,nq,heads] f32 score tensor materialised three times) for the UNet's
// self/cross attention (mlblock_nn.c:190-231) and the VAE's spatial attention (vae.c:46-74).
//
// One CTA owns TWO tiles of 128 query rows of one (head, image) and streams the keys in blocks of
// 128, ping-ponging the tiles so the tensor core works on one while the other is in softmax:
//   warp 0        TMA producer: both Q tiles once, then a ring of K and V tiles (4-D tensor maps
//                 straight over the token-major [token][head][d] activations: the reference's head
//                 split/merge copies, mlblock_nn.c:204-226, do not exist)
//   warp 1        MMA issuer (one thread): S_t = Q_t K^T into the tile's TMEM score buffer
//                 (tcgen05.mma, both operands K-major from shared memory); O_t,j = P_t V with the
//                 A operand P read FROM TENSOR MEMORY (the softmax warps overwrite the consumed part
//                 of S_t with the f16 probabilities) and V as an MN-major B operand, i.e. in its
//                 natural [key][d] layout
//   warps 2..5 / 6..9   softmax + accumulate for tile a / b, one query row per thread: tcgen05.ld of
//                 the scores, running max / sum in the exp2 domain (ex2.approx), P written back with
//                 tcgen05.st; the PV tile of the previous block is folded into the register
//                 accumulator with the online-softmax correction while the tensor core is busy.
// Head dims 40 / 80 (SD1.x) are handled by TMA zero-fill up to the next multiple of 16.
// Roofline: tensor (4*nq*nk*d FLOP per head); for d <= 64 the MUFU.EX2 rate (16/clk/SM) is the
// practical limit (1024 clk per 128x128 tile vs 2*(d/16)*64 clk of MMA).
#include "kernels.h"
#include "tc_common.cuh"
#include <algorithm>

namespace b200 {

constexpr int AQ = 128;      // queries per tile (two tiles per CTA)
constexpr int AK = 128;      // keys per block
constexpr int ACH = 64;      // columns per TMA chunk (128 B)
constexpr int CHUNK_BYTES = 128 * 128;   // [128 rows][64 f16]
constexpr int A_TMEM_COLS = 512;
constexpr int A_MAX_STAGES = 4;

struct AttnParams {
	int d, d16, dchunks, nq, nk, H, B, stages, nblk;
	float scale_log2;
	void* o; long long so_t, so_h, so_b;
};

struct AttnTC {
	CUtensorMap tmQ, tmK, tmV;
	AttnParams p;
	dim3 grid;
	size_t smem;
};

__device__ __forceinline__ float ex2_approx(float x)
{ float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r)
{
	asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
		:: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
		   "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem desc]   (A operand from tensor memory)
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
	asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
		"tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
		:: "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

template <int D16MAX, int NT>     // NT = query tiles per CTA (2: ping-pong, 320 threads; 1: wide heads, 192 threads)
__global__ void __launch_bounds__(64 + 128 * NT, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
	const AttnParams p)
{
	extern __shared__ __align__(1024) uint8_t smem_raw[];
	uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
	const int tile_bytes = p.dchunks * CHUNK_BYTES;
	uint8_t* sQ = smem;                                   // [2 tiles]
	uint8_t* sK = sQ + NT * tile_bytes;                   // [stages]
	uint8_t* sV = sK + (size_t)p.stages * tile_bytes;     // [stages]
	uint64_t* bars = (uint64_t*)(sV + (size_t)p.stages * tile_bytes);
	uint64_t* q_full = bars;                        // [2]
	uint64_t* k_full = q_full + 2;                  // [stages]
	uint64_t* k_empty = k_full + A_MAX_STAGES;
	uint64_t* v_full = k_empty + A_MAX_STAGES;
	uint64_t* v_empty = v_full + A_MAX_STAGES;
	uint64_t* s_full = v_empty + A_MAX_STAGES;      // [2 tiles]  QK done
	uint64_t* p_full = s_full + 2;                  // [2 tiles]  probabilities written (128 arrivals)
	uint64_t* pv_full = p_full + 2;                 // [2 tiles]  PV done
	uint64_t* pv_empty = pv_full + 2;               // [2 tiles]  PV tile consumed (128 arrivals)
	uint32_t* tmem_slot = (uint32_t*)(pv_empty + 2);

	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int q0 = blockIdx.x * (NT * AQ), h = blockIdx.y, b = blockIdx.z;
	const int ntile = (NT == 2 && q0 + AQ < p.nq) ? 2 : 1;     // the second tile may be entirely out of range

	if (threadIdx.x == 0) {
		tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
		for (int s = 0; s < p.stages; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1); }
		for (int t = 0; t < 2; ++t) { mbar_init(&q_full[t], 1); mbar_init(&s_full[t], 1); mbar_init(&p_full[t], 128); mbar_init(&pv_full[t], 1); mbar_init(&pv_empty[t], 128); }
		fence_barrier_init();
	}
	if (warp == 1) tmem_alloc(tmem_slot, A_TMEM_COLS);
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem_base = *tmem_slot;
	// TMEM columns: S_a [0,128) (P_a aliases [0,64)), S_b [128,256), O_a tile [256,384), O_b tile [384,512)

	if (warp == 0) {
		if (lane == 0) {
			for (int t = 0; t < ntile; ++t) {
				mbar_expect_tx(&q_full[t], tile_bytes);
				for (int c = 0; c < p.dchunks; ++c) tma_load_4d(sQ + t * tile_bytes + c * CHUNK_BYTES, &tmQ, &q_full[t], c * ACH, q0 + t * AQ, h, b);
			}
			for (int j = 0; j < p.nblk; ++j) {
				const int s = j % p.stages; const uint32_t ph = (uint32_t)(j / p.stages) & 1;
				mbar_wait(&k_empty[s], ph ^ 1);
				mbar_expect_tx(&k_full[s], tile_bytes);
				for (int c = 0; c < p.dchunks; ++c) tma_load_4d(sK + (size_t)s * tile_bytes + c * CHUNK_BYTES, &tmK, &k_full[s], c * ACH, j * AK, h, b);
				mbar_wait(&v_empty[s], ph ^ 1);
				mbar_expect_tx(&v_full[s], tile_bytes);
				for (int c = 0; c < p.dchunks; ++c) tma_load_4d(sV + (size_t)s * tile_bytes + c * CHUNK_BYTES, &tmV, &v_full[s], c * ACH, j * AK, h, b);
			}
		}
	} else if (warp == 1) {
		if (lane == 0) {
			const uint32_t idesc_qk = make_idesc_f16(AQ, AK, 0, 0);
			const uint32_t idesc_pv = make_idesc_f16(AQ, p.d16, 0, 1);     // A = P (TMEM), B = V MN-major ([key][d], d contiguous)
			auto issue_qk = [&](int t, int j) {
				const int s = j % p.stages;
				if (t == 0) { mbar_wait(&k_full[s], (uint32_t)(j / p.stages) & 1); tc_fence_after(); }
				const uint32_t q_addr = smem_u32(sQ + t * tile_bytes);
				const uint32_t k_addr = smem_u32(sK + (size_t)s * tile_bytes);
				int first = 1;
				for (int c = 0; c < p.dchunks; ++c)
					for (int kk = 0; kk < 4; ++kk) {
						if (c * ACH + kk * 16 >= p.d16) break;
						uint64_t ad = make_smem_desc_sw128(q_addr + c * CHUNK_BYTES + kk * 32, 16, 1024);
						uint64_t bd = make_smem_desc_sw128(k_addr + c * CHUNK_BYTES + kk * 32, 16, 1024);
						umma_f16(tmem_base + t * 128, ad, bd, idesc_qk, first ? 0u : 1u);
						first = 0;
					}
				umma_commit(&s_full[t]);
				if (t == ntile - 1) umma_commit(&k_empty[s]);
			};
			auto issue_pv = [&](int t, int j) {
				const int s = j % p.stages;
				if (t == 0) mbar_wait(&v_full[s], (uint32_t)(j / p.stages) & 1);
				mbar_wait(&p_full[t], (uint32_t)j & 1);
				mbar_wait(&pv_empty[t], ((uint32_t)j & 1) ^ 1);
				tc_fence_after();
				const uint32_t v_addr = smem_u32(sV + (size_t)s * tile_bytes);
				const int valid = min(AK, p.nk - j * AK);
				int first = 1;
				for (int kk = 0; kk < AK / 16; ++kk) {
					const int key = kk * 16;
					if (key >= valid) break;
					// V tile: rows = keys (128 B each), 8-row swizzle atoms 1024 B apart (SBO), next 64 d-columns CHUNK_BYTES away (LBO)
					uint64_t bd = make_smem_desc_sw128(v_addr + key * 128, CHUNK_BYTES, 1024);
					// P in tensor memory: 16 keys = 8 packed 32-bit columns per k-step
					umma_f16_ts(tmem_base + 256 + t * 128, tmem_base + t * 128 + kk * 8, bd, idesc_pv, first ? 0u : 1u);
					first = 0;
				}
				umma_commit(&pv_full[t]);
				if (t == ntile - 1) umma_commit(&v_empty[s]);
			};
			for (int t = 0; t < ntile; ++t) { mbar_wait(&q_full[t], 0); }
			for (int t = 0; t < ntile; ++t) issue_qk(t, 0);
			for (int j = 0; j < p.nblk; ++j) {
				for (int t = 0; t < ntile; ++t) {
					issue_pv(t, j);                             // in-order after it: S_t/P_t may be overwritten
					if (j + 1 < p.nblk) issue_qk(t, j + 1);
				}
			}
		}
	} else {
		// ===== softmax + accumulate: group 0 = warps 2..5 (tile a), group 1 = warps 6..9 (tile b) =====
		const int t = (warp - 2) >> 2;
		if (t < ntile) {
			const int quarter = warp & 3;
			const int r = quarter * 32 + lane;
			const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
			const uint32_t ts = tmem_base + t * 128 + lane_off;           // scores / probabilities of this row
			const uint32_t tpv = tmem_base + 256 + t * 128 + lane_off;    // PV tile of this row
			const float sl2 = p.scale_log2;
			float m = -INFINITY, l = 0.f, corr_saved = 0.f;
			float acc[D16MAX];
			#pragma unroll
			for (int i = 0; i < D16MAX; ++i) acc[i] = 0.f;

			auto accumulate = [&](int jj, float corr) {
				mbar_wait(&pv_full[t], (uint32_t)jj & 1);
				tc_fence_after();
				#pragma unroll
				for (int c0 = 0; c0 < D16MAX; c0 += 16) {
					if (c0 < p.d16) {
						uint32_t v[16];
						tmem_ld16(tpv + c0, v);
						tmem_ld_wait();
						#pragma unroll
						for (int i = 0; i < 16; ++i) acc[c0 + i] = fmaf(acc[c0 + i], corr, __uint_as_float(v[i]));
					}
				}
				tc_fence_before();
				mbar_arrive(&pv_empty[t]);
			};

			for (int j = 0; j < p.nblk; ++j) {
				const int valid = p.nk - j * AK;      // >= 128 for all but a partial last block
				mbar_wait(&s_full[t], (uint32_t)j & 1);
				tc_fence_after();
				// Two passes over the 128 scores of this row (max, then probabilities), 32 columns at a time,
				// with the next tcgen05.ld always in flight while the current chunk is processed. Four
				// independent max / sum chains keep the FP pipes busy instead of one serial dependency.
				uint32_t va[32], vb[32];
				float mx4[4] = { -INFINITY, -INFINITY, -INFINITY, -INFINITY };
				float rs4[4] = { 0.f, 0.f, 0.f, 0.f };
				float m_new = 0.f;
				const bool full = valid >= AK;
				auto maxupd = [&](const uint32_t* v, int c) {
					if (full) {
						#pragma unroll
						for (int i = 0; i < 32; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(v[i]));
					} else {
						#pragma unroll
						for (int i = 0; i < 32; ++i) if (c * 32 + i < valid) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(v[i]));
					}
				};
				auto proc = [&](const uint32_t* v, int c) {
					uint32_t packed[16];
					#pragma unroll
					for (int i = 0; i < 32; i += 2) {
						float p0 = ex2_approx(fmaf(__uint_as_float(v[i]), sl2, -m_new));
						float p1 = ex2_approx(fmaf(__uint_as_float(v[i + 1]), sl2, -m_new));
						if (!full) { if (c * 32 + i >= valid) p0 = 0.f; if (c * 32 + i + 1 >= valid) p1 = 0.f; }
						rs4[(i >> 1) & 3] += p0 + p1;
						__half2 hh = __floats2half2_rn(p0, p1);
						packed[i >> 1] = *reinterpret_cast<uint32_t*>(&hh);
					}
					tmem_st16(ts + c * 16, packed);    // columns [16c, 16c+16) lie inside the already consumed scores
				};
				tmem_ld32(ts, va);
				tmem_ld_wait(); tmem_ld32(ts + 32, vb); maxupd(va, 0);
				tmem_ld_wait(); tmem_ld32(ts + 64, va); maxupd(vb, 1);
				tmem_ld_wait(); tmem_ld32(ts + 96, vb); maxupd(va, 2);
				tmem_ld_wait(); tmem_ld32(ts, va);      maxupd(vb, 3);
				const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
				m_new = fmaxf(m, mx * sl2);
				const float corr = ex2_approx(m - m_new);      // m = -inf on the first block -> 0
				tmem_ld_wait(); tmem_ld32(ts + 32, vb); proc(va, 0);
				tmem_ld_wait(); tmem_ld32(ts + 64, va); proc(vb, 1);
				tmem_ld_wait(); tmem_ld32(ts + 96, vb); proc(va, 2);
				tmem_ld_wait();                         proc(vb, 3);
				const float rowsum = (rs4[0] + rs4[1]) + (rs4[2] + rs4[3]);
				tmem_st_wait();
				tc_fence_before();
				mbar_arrive(&p_full[t]);
				l = fmaf(l, corr, rowsum);
				m = m_new;
				if (j > 0) accumulate(j - 1, corr_saved);
				corr_saved = corr;
			}
			accumulate(p.nblk - 1, corr_saved);

			const long long tok = (long long)q0 + t * AQ + r;
			if (tok < p.nq) {
				const float inv = l > 0.f ? 1.0f / l : 0.f;
				__half* op = (__half*)p.o + tok * p.so_t + (long long)h * p.so_h + (long long)b * p.so_b;
				const bool vec = ((((uintptr_t)op) & 15) == 0);
				#pragma unroll
				for (int c0 = 0; c0 < D16MAX; c0 += 8) {
					if (c0 < p.d) {
						if (vec && c0 + 8 <= p.d) {
							uint4 o4; __half2* hp = reinterpret_cast<__half2*>(&o4);
							#pragma unroll
							for (int i = 0; i < 4; ++i) hp[i] = __floats2half2_rn(acc[c0 + 2 * i] * inv, acc[c0 + 2 * i + 1] * inv);
							*reinterpret_cast<uint4*>(op + c0) = o4;
						} else {
							#pragma unroll
							for (int i = 0; i < 8; ++i) if (c0 + i < p.d) op[c0 + i] = __float2half_rn(acc[c0 + i] * inv);
						}
					}
				}
			}
		}
	}
	tc_fence_before();
	__syncthreads();
	if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, A_TMEM_COLS); }
}

// ------------------------------------------------------------------ host
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
	const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static bool encode4(CUtensorMap* tm, const void* base, int64_t d, int64_t n, int64_t H, int64_t B, int64_t st_t, int64_t st_h, int64_t st_b)
{
	static PFN_encodeTiled fn = nullptr;
	if (!fn) {
		void* ptr = nullptr; cudaDriverEntryPointQueryResult qres;
		CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres));
		fn = (PFN_encodeTiled)ptr;
	}
	// degenerate dims get a harmless stride
	if (H == 1) st_h = st_t * n;
	if (B == 1) st_b = st_h * H;
	cuuint64_t dims[4] = { (cuuint64_t)d, (cuuint64_t)n, (cuuint64_t)H, (cuuint64_t)B };
	cuuint64_t strides[3] = { (cuuint64_t)st_t * 2, (cuuint64_t)st_h * 2, (cuuint64_t)st_b * 2 };
	cuuint32_t box[4] = { ACH, 128, 1, 1 }, es[4] = { 1, 1, 1, 1 };
	CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, es,
		CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	return r == CUDA_SUCCESS;
}

static bool operand_ok(const View& v, int dim_d, int dim_t)
{
	if (v.dt != DT_F16 || v.st[dim_d] != 1 || ((uintptr_t)v.ptr & 15)) return false;
	if (v.st[dim_t] % 8 || (v.ne[2] > 1 && v.st[2] % 8) || (v.ne[3] > 1 && v.st[3] % 8)) return false;
	return true;
}

bool attn_tc_supported(const View& o, const View& q, const View& k, const View& v, bool causal)
{
	if (causal) return false;
	int d = (int)q.ne[0];
	if (d > 160 || d % 8) return false;
	if (!operand_ok(q, 0, 1) || !operand_ok(k, 0, 1) || !operand_ok(v, 1, 0)) return false;
	if (o.dt != DT_F16 || o.st[0] != 1) return false;
	return true;
}

AttnTC* attn_tc_prepare(const View& o, const View& q, const View& k, const View& v, float scale)
{
	AttnTC* a = new AttnTC();
	AttnParams& p = a->p;
	p.d = (int)q.ne[0]; p.d16 = (p.d + 15) / 16 * 16; p.dchunks = (p.d + ACH - 1) / ACH;
	p.nq = (int)q.ne[1]; p.nk = (int)k.ne[1]; p.H = (int)q.ne[2]; p.B = (int)q.ne[3];
	p.nblk = (p.nk + AK - 1) / AK;
	p.scale_log2 = scale * 1.4426950408889634f;
	p.o = o.ptr; p.so_t = o.st[1]; p.so_h = o.st[2]; p.so_b = o.st[3];
	const size_t tile = (size_t)p.dchunks * CHUNK_BYTES;
	p.stages = std::max(1, std::min(p.nblk, A_MAX_STAGES));
	const int nt = p.d16 > 128 ? 1 : 2;
	auto total = [&]() { return tile * (nt + 2 * p.stages) + 1024 + 512; };
	while (total() > 220 * 1024 && p.stages > 1) p.stages--;
	a->smem = total();
	a->grid = dim3((unsigned)((p.nq + nt * AQ - 1) / (nt * AQ)), (unsigned)p.H, (unsigned)p.B);
	bool ok = encode4(&a->tmQ, q.ptr, p.d, p.nq, p.H, p.B, q.st[1], q.st[2], q.st[3]) &&
	          encode4(&a->tmK, k.ptr, p.d, p.nk, p.H, p.B, k.st[1], k.st[2], k.st[3]) &&
	          encode4(&a->tmV, v.ptr, p.d, p.nk, p.H, p.B, v.st[0], v.st[2], v.st[3]);   // v is the [nk, d, H, B] view: token stride = st[0]
	if (!ok) { delete a; return nullptr; }
	return a;
}

void attn_tc_launch(cudaStream_t s, AttnTC* a)
{
	static bool attr_set = false;
	if (!attr_set) {
		CUDA_CHECK(cudaFuncSetAttribute(attn_tc_kernel<64, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
		CUDA_CHECK(cudaFuncSetAttribute(attn_tc_kernel<128, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
		CUDA_CHECK(cudaFuncSetAttribute(attn_tc_kernel<160, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
		attr_set = true;
	}
	if (a->p.d16 <= 64) attn_tc_kernel<64, 2><<<a->grid, 320, a->smem, s>>>(a->tmQ, a->tmK, a->tmV, a->p);
	else if (a->p.d16 <= 128) attn_tc_kernel<128, 2><<<a->grid, 320, a->smem, s>>>(a->tmQ, a->tmK, a->tmV, a->p);
	else attn_tc_kernel<160, 1><<<a->grid, 192, a->smem, s>>>(a->tmQ, a->tmK, a->tmV, a->p);
	g_stats.kernel_launches++;
}

void attn_tc_free(AttnTC* a) { delete a; }

}  // namespace b200
