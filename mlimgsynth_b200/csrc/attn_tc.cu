// attn_tc.cu -- flash-style fused attention on the 5th-gen tensor cores (sm_100a).
// Replaces the reference's unfused ggml_nn_attention chain (ggml_extend.c:200-221: QK^T, scale,
// softmax, PV with the [nk,nq,heads] f32 score tensor materialised three times) for the UNet's
// self/cross attention (mlblock_nn.c:190-231) and the VAE's spatial attention (vae.c:46-74).
//
// One CTA owns 128 query rows of one (head, image) and streams the keys in blocks of 128:
//   warp 0      TMA producer: Q once, then a ring of K and V tiles (4-D tensor maps straight over the
//               token-major [token][head][d] activations: no head split/merge copies exist)
//   warp 1      MMA issuer: S = Q K^T (tcgen05.mma, K-major operands) into one of two TMEM score
//               buffers, then PV = P V (A = P from shared memory, B = V as an MN-major operand, so V
//               is consumed in its natural [key][d] layout) into a TMEM output tile
//   warps 2..5  softmax + accumulate, one query row per thread: tcgen05.ld of the scores, running
//               max / sum in the exp2 domain, P written as f16 into the 128B-swizzled K-major tile
//               the second MMA reads; the PV tile of the previous block is added to the register
//               accumulator with the usual online-softmax correction while the tensor core works on
//               the next block.
// Head dims 40/80 (SD1.x) are handled by TMA zero-fill up to the next multiple of 16.
#include "kernels.h"
#include "tc_common.cuh"
#include <algorithm>

namespace b200 {

constexpr int AQ = 128;      // queries per CTA
constexpr int AK = 128;      // keys per block
constexpr int ACH = 64;      // columns per TMA chunk (128 B)
constexpr int CHUNK_BYTES = 128 * 128;   // [128 rows][64 f16]
constexpr int A_TMEM_COLS = 512;
constexpr int A_MAX_STAGES = 3;

struct AttnParams {
	int d, d16, dchunks, nq, nk, H, B, stages, pbufs, nblk;
	float scale_log2;
	void* o; long long so_t, so_h, so_b;
};

struct AttnTC {
	CUtensorMap tmQ, tmK, tmV;
	AttnParams p;
	dim3 grid;
	size_t smem;
};

template <int D16MAX>
__global__ void __launch_bounds__(192, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
	const AttnParams p)
{
	extern __shared__ __align__(1024) uint8_t smem_raw[];
	uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
	const int tile_bytes = p.dchunks * CHUNK_BYTES;
	uint8_t* sQ = smem;
	uint8_t* sK = sQ + tile_bytes;
	uint8_t* sV = sK + (size_t)p.stages * tile_bytes;
	uint8_t* sP = sV + (size_t)p.stages * tile_bytes;
	uint64_t* bars = (uint64_t*)(sP + (size_t)p.pbufs * 2 * CHUNK_BYTES);
	uint64_t* q_full = bars;
	uint64_t* k_full = bars + 1;                    // [stages]
	uint64_t* k_empty = k_full + A_MAX_STAGES;
	uint64_t* v_full = k_empty + A_MAX_STAGES;
	uint64_t* v_empty = v_full + A_MAX_STAGES;
	uint64_t* s_full = v_empty + A_MAX_STAGES;      // [2]
	uint64_t* s_empty = s_full + 2;
	uint64_t* p_full = s_empty + 2;                 // [2]
	uint64_t* p_empty = p_full + 2;
	uint64_t* pv_full = p_empty + 2;
	uint64_t* pv_empty = pv_full + 1;
	uint32_t* tmem_slot = (uint32_t*)(pv_empty + 1);

	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int q0 = blockIdx.x * AQ, h = blockIdx.y, b = blockIdx.z;

	if (threadIdx.x == 0) {
		tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
		mbar_init(q_full, 1);
		for (int s = 0; s < p.stages; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1); }
		for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 128); mbar_init(&p_full[i], 128); mbar_init(&p_empty[i], 1); }
		mbar_init(pv_full, 1); mbar_init(pv_empty, 128);
		fence_barrier_init();
	}
	if (warp == 1) tmem_alloc(tmem_slot, A_TMEM_COLS);
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem_base = *tmem_slot;
	const uint32_t tm_S[2] = { tmem_base, tmem_base + 128 };
	const uint32_t tm_PV = tmem_base + 256;

	if (warp == 0) {
		if (lane == 0) {
			mbar_expect_tx(q_full, tile_bytes);
			for (int c = 0; c < p.dchunks; ++c) tma_load_4d(sQ + c * CHUNK_BYTES, &tmQ, q_full, c * ACH, q0, h, b);
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
			const uint32_t idesc_pv = make_idesc_f16(AQ, p.d16, 0, 1);     // B = V, MN-major ([key][d], d contiguous)
			const uint32_t q_addr = smem_u32(sQ);
			auto issue_qk = [&](int j) {
				const int s = j % p.stages;
				mbar_wait(&k_full[s], (uint32_t)(j / p.stages) & 1);
				mbar_wait(&s_empty[j & 1], ((uint32_t)(j >> 1) & 1) ^ 1);
				tc_fence_after();
				const uint32_t k_addr = smem_u32(sK + (size_t)s * tile_bytes);
				int first = 1;
				for (int c = 0; c < p.dchunks; ++c)
					for (int kk = 0; kk < 4; ++kk) {
						if (c * ACH + kk * 16 >= p.d16) break;
						uint64_t ad = make_smem_desc_sw128(q_addr + c * CHUNK_BYTES + kk * 32, 16, 1024);
						uint64_t bd = make_smem_desc_sw128(k_addr + c * CHUNK_BYTES + kk * 32, 16, 1024);
						umma_f16(tm_S[j & 1], ad, bd, idesc_qk, first ? 0u : 1u);
						first = 0;
					}
				umma_commit(&s_full[j & 1]);
				umma_commit(&k_empty[s]);
			};
			mbar_wait(q_full, 0);
			issue_qk(0);
			for (int j = 0; j < p.nblk; ++j) {
				if (j + 1 < p.nblk) issue_qk(j + 1);
				const int s = j % p.stages, pb = j % p.pbufs;
				mbar_wait(&v_full[s], (uint32_t)(j / p.stages) & 1);
				mbar_wait(&p_full[pb], (uint32_t)(j / p.pbufs) & 1);
				mbar_wait(pv_empty, ((uint32_t)j & 1) ^ 1);
				tc_fence_after();
				const uint32_t v_addr = smem_u32(sV + (size_t)s * tile_bytes);
				const uint32_t p_addr = smem_u32(sP + (size_t)pb * 2 * CHUNK_BYTES);
				const int valid = min(AK, p.nk - j * AK);
				int first = 1;
				for (int kc = 0; kc < 2; ++kc)
					for (int kk = 0; kk < 4; ++kk) {
						const int key = kc * ACH + kk * 16;
						if (key >= valid) break;
						uint64_t ad = make_smem_desc_sw128(p_addr + kc * CHUNK_BYTES + kk * 32, 16, 1024);
						// V tile: rows = keys (128 B each), 8-row swizzle atoms 1024 B apart (SBO), next 64 d-columns CHUNK_BYTES away (LBO)
						uint64_t bd = make_smem_desc_sw128(v_addr + key * 128, CHUNK_BYTES, 1024);
						umma_f16(tm_PV, ad, bd, idesc_pv, first ? 0u : 1u);
						first = 0;
					}
				umma_commit(pv_full);
				umma_commit(&v_empty[s]);
				umma_commit(&p_empty[pb]);
			}
		}
	} else {
		// ===== softmax + accumulate: one query row per thread =====
		const int quarter = warp & 3;
		const int r = quarter * 32 + lane;
		const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
		float m = -INFINITY, l = 0.f, corr_saved = 0.f;
		float acc[D16MAX];
		#pragma unroll
		for (int i = 0; i < D16MAX; ++i) acc[i] = 0.f;

		auto accumulate = [&](int jj, float corr) {
			mbar_wait(pv_full, (uint32_t)jj & 1);
			tc_fence_after();
			#pragma unroll
			for (int c0 = 0; c0 < D16MAX; c0 += 16) {
				if (c0 < p.d16) {
					uint32_t v[16];
					tmem_ld16(tm_PV + lane_off + c0, v);
					tmem_ld_wait();
					#pragma unroll
					for (int i = 0; i < 16; ++i) acc[c0 + i] = acc[c0 + i] * corr + __uint_as_float(v[i]);
				}
			}
			tc_fence_before();
			mbar_arrive(pv_empty);
		};

		for (int j = 0; j < p.nblk; ++j) {
			const int valid = min(AK, p.nk - j * AK);
			mbar_wait(&s_full[j & 1], (uint32_t)(j >> 1) & 1);
			tc_fence_after();
			const uint32_t ts = tm_S[j & 1] + lane_off;
			// pass A: row maximum of the valid keys
			float mx = -INFINITY;
			#pragma unroll
			for (int c = 0; c < 4; ++c) {
				uint32_t v[32];
				tmem_ld32(ts + c * 32, v);
				tmem_ld_wait();
				#pragma unroll
				for (int i = 0; i < 32; ++i) if (c * 32 + i < valid) mx = fmaxf(mx, __uint_as_float(v[i]));
			}
			const float m_new = fmaxf(m, mx * p.scale_log2);
			const float corr = (m == -INFINITY) ? 0.f : exp2f(m - m_new);
			// the P tile must have been consumed by the PV MMA that last used this buffer
			const int pb = j % p.pbufs;
			mbar_wait(&p_empty[pb], ((uint32_t)(j / p.pbufs) & 1) ^ 1);
			uint8_t* prow = sP + (size_t)pb * 2 * CHUNK_BYTES + r * 128;
			float rowsum = 0.f;
			#pragma unroll
			for (int c = 0; c < 4; ++c) {
				uint32_t v[32];
				tmem_ld32(ts + c * 32, v);
				tmem_ld_wait();
				uint32_t packed[16];
				#pragma unroll
				for (int i = 0; i < 32; i += 2) {
					float p0 = (c * 32 + i < valid) ? exp2f(__uint_as_float(v[i]) * p.scale_log2 - m_new) : 0.f;
					float p1 = (c * 32 + i + 1 < valid) ? exp2f(__uint_as_float(v[i + 1]) * p.scale_log2 - m_new) : 0.f;
					__half2 hh = __floats2half2_rn(p0, p1);
					// sum what the tensor core will actually multiply (the f16-rounded probabilities)
					float2 back = __half22float2(hh);
					rowsum += back.x + back.y;
					packed[i >> 1] = *reinterpret_cast<uint32_t*>(&hh);
				}
				// 32 keys = 64 B = four 16-byte chunks of this row; key chunk kc = c / 2, chunk index inside the 128 B row = (c & 1) * 4 + q
				uint8_t* base = prow + (c >> 1) * CHUNK_BYTES;
				#pragma unroll
				for (int q = 0; q < 4; ++q) {
					const int ci = ((c & 1) * 4 + q) ^ (r & 7);       // 128B swizzle: chunk index XOR (row mod 8)
					*reinterpret_cast<uint4*>(base + ci * 16) = make_uint4(packed[q * 4], packed[q * 4 + 1], packed[q * 4 + 2], packed[q * 4 + 3]);
				}
			}
			tc_fence_before();
			mbar_arrive(&s_empty[j & 1]);          // score buffer may be overwritten by QK(j+2)
			fence_proxy_async();                   // generic-proxy smem writes -> visible to the tensor-core (async) proxy
			mbar_arrive(&p_full[pb]);
			l = l * corr + rowsum;
			m = m_new;
			if (j > 0) accumulate(j - 1, corr_saved);
			corr_saved = corr;
		}
		accumulate(p.nblk - 1, corr_saved);

		const long long t = (long long)q0 + r;
		if (t < p.nq) {
			const float inv = l > 0.f ? 1.0f / l : 0.f;
			__half* op = (__half*)p.o + t * p.so_t + (long long)h * p.so_h + (long long)b * p.so_b;
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
	if (d > 128 || d % 8) return false;
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
	p.pbufs = 2;
	p.stages = p.nblk >= 3 ? 3 : std::max(1, p.nblk);
	auto total = [&]() { return tile * (1 + 2 * p.stages) + (size_t)p.pbufs * 2 * CHUNK_BYTES + 1024 + 256; };
	while (total() > 220 * 1024 && p.stages > 2) p.stages--;
	if (total() > 220 * 1024) p.pbufs = 1;
	a->smem = total();
	a->grid = dim3((unsigned)((p.nq + AQ - 1) / AQ), (unsigned)p.H, (unsigned)p.B);
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
		CUDA_CHECK(cudaFuncSetAttribute(attn_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
		CUDA_CHECK(cudaFuncSetAttribute(attn_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
		attr_set = true;
	}
	if (a->p.d16 <= 64) attn_tc_kernel<64><<<a->grid, 192, a->smem, s>>>(a->tmQ, a->tmK, a->tmV, a->p);
	else attn_tc_kernel<128><<<a->grid, 192, a->smem, s>>>(a->tmQ, a->tmK, a->tmV, a->p);
	g_stats.kernel_launches++;
}

void attn_tc_free(AttnTC* a) { delete a; }

}  // namespace b200
