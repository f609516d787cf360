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
// Kernels in this file, by head width and context length (attn_tc_launch picks; every variant is selectable by environment):
//   attn_ap_kernel     heads <= 48 wide, more than one key block: two query tiles per CTA, scores three blocks deep (SD1.x level 0)
//   attn_split_kernel  heads 49..64 wide: key halves, eight softmax warps per tile, two CTAs per SM (SDXL, SD2.x)
//   attn_tc_kernel     wider heads (80, 160) and the round-1 forms; attn_split4_kernel: measured-slower variants kept for reference
//   attn_kv1_kernel    one key block (the 77-token text context): a CTA walks the query tiles of its (head, image)
// Roofline: tensor (4*nq*nk*d FLOP per head); for d <= 64 the MUFU.EX2 rate (16/clk/SM) is the
// practical limit (1024 clk per 128x128 tile vs 2*(d/16)*64 clk of MMA).
#include "kernels.h"
#include "tc_common.cuh"
#include <algorithm>
#include <type_traits>

namespace b200 {

constexpr int AQ = 128;      // queries per tile (two tiles per CTA)
constexpr int AK = 128;      // keys per block
constexpr int ACH = 64;      // columns per TMA chunk (128 B)
constexpr int CHUNK_BYTES = 128 * 128;   // [128 rows][64 f16]
constexpr int A_TMEM_COLS = 512;
constexpr int A_MAX_STAGES = 4;

struct AttnParams {
	int d, d16, dchunks, nq, nk, H, B, stages, nblk, pingpong, sep_p, dual, npoly, split, park, pk;
	float scale_log2;
	void* o; long long so_t, so_h, so_b;
	long long* trace;     // debug timeline (tools/attn_trace.cu); null in production
};
// Timeline events of CTA (0,0,0): slot = role * 64 * 8 + j * 8 + ev, roles 0/1 = softmax warp of tile a/b (lane 0 of its
// first warp), 2 = the MMA thread.
#define ATTN_TR(role, j, ev) do { if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (j) < 64) \
	p.trace[((role) * 64 + (j)) * 8 + (ev)] = clock64(); } while (0)

struct AttnTC {
	CUtensorMap tmQ, tmK, tmV;
	AttnParams p;
	dim3 grid;
	size_t smem;
	bool kv1 = false;     // single key block: attn_kv1_kernel
};

__device__ __forceinline__ float ex2_approx(float x)
{ float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// 2^x for x <= ~0 on the FMA / ALU pipes: x = n + f, n = round(x), f in [-0.5, 0.5]; 2^f by a degree-3 minimax polynomial
// (relative error 7.5e-5, below f16 rounding); n is added to the exponent field (the low mantissa bits of x + 1.5 * 2^23).
__device__ __forceinline__ float ex2_poly(float x)
{
	x = fmaxf(x, -126.0f);
	const float magic = 12582912.0f;
	const float xr = x + magic;
	const float n = xr - magic;
	const float f = x - n;
	float p = fmaf(f, 0.0551716648f, 0.242611125f);
	p = fmaf(p, f, 0.693260968f);
	p = fmaf(p, f, 0.999928057f);
	return __int_as_float(__float_as_int(p) + (__float_as_int(xr) << 23));
}

__device__ __forceinline__ float max3f(float a, float b, float c)
{ float y; asm("max.f32 %0, %1, %2, %3;" : "=f"(y) : "f"(a), "f"(b), "f"(c)); return y; }

// Packed f32x2 arithmetic (sm_100: FFMA2 / FADD2 take ONE issue slot for two elements). The softmax of narrow heads is bound by
// issue slots as much as by the MUFU pipe (profiles/r2_attention_microbench.md), so the scale-and-subtract, the row sums and the
// polynomial exponentials run on register pairs.
__device__ __forceinline__ uint64_t pk2(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ uint64_t pk2u(uint32_t a, uint32_t b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ void unpk2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c)
{ uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) { uint64_t r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

// ex2_poly on a register pair: the clamp and the exponent insertion stay scalar (no packed min/max or integer forms), the
// rounding, the reduction and the polynomial are packed: 10 issue slots per pair against 2 MUFU.EX2 (16 MUFU cycles per warp).
__device__ __forceinline__ void ex2_poly2(uint64_t x2, float& e0, float& e1)
{
	float x0, x1; unpk2(x2, x0, x1);
	x0 = fmaxf(x0, -126.0f); x1 = fmaxf(x1, -126.0f);
	const uint64_t xc = pk2(x0, x1);
	const uint64_t xr = add2(xc, pk2(12582912.0f, 12582912.0f));
	const uint64_t n = sub2(xr, pk2(12582912.0f, 12582912.0f));
	const uint64_t f = sub2(xc, n);
	uint64_t q = fma2(f, pk2(0.0551716648f, 0.0551716648f), pk2(0.242611125f, 0.242611125f));
	q = fma2(q, f, pk2(0.693260968f, 0.693260968f));
	q = fma2(q, f, pk2(0.999928057f, 0.999928057f));
	float q0, q1, r0, r1; unpk2(q, q0, q1); unpk2(xr, r0, r1);
	e0 = __int_as_float(__float_as_int(q0) + (__float_as_int(r0) << 23));
	e1 = __int_as_float(__float_as_int(q1) + (__float_as_int(r1) << 23));
}

__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" :: "r"(id), "r"(n) : "memory"); }

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

// NT = query tiles per CTA (2: ping-pong, 320 threads; 1: 192 threads). <64, 1> is the "dual" form: one tile, 256 tensor
// memory columns and <= 113 KB of shared memory, so that TWO CTAs share an SM -- two softmax warps per SM sub-partition
// keep the MUFU unit busier than one (8.6 vs 9.6 clk per exponential, tools/pipe_rates.cu) and each CTA's QK/PV latency
// hides behind the other's exponentials without an explicit ping-pong.
template <int D16MAX, int NT, int N_POLY = 0>       // N_POLY of every 8 exponentials run on the FMA pipe instead of the MUFU unit
__global__ void __launch_bounds__(64 + 128 * NT, (NT == 1 && D16MAX <= 64) ? 2 : 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
	const AttnParams p)
{
	extern __shared__ __align__(1024) uint8_t smem_raw[];
	uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
	const int tile_bytes = p.dchunks * CHUNK_BYTES;
	uint8_t* sQ = smem;                                   // [NT tiles]
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
	uint64_t* s_free = pv_full + 2;                 // [2 tiles]  scores of the block are in registers (one arrival per softmax warp)
	uint32_t* tmem_slot = (uint32_t*)(s_free + 2);

	const int warp = warp_id_uniform(), lane = threadIdx.x & 31;
	const int q0 = blockIdx.x * (NT * AQ), h = blockIdx.y, b = blockIdx.z;
	const int ntile = (NT == 2 && q0 + AQ < p.nq) ? 2 : 1;     // the second tile may be entirely out of range
	// TMEM columns: S_t at t*128, O_t behind them, then -- if there is room -- P_t (64 columns of packed f16 pairs);
	// otherwise P_t aliases the first 64 columns of S_t.
	constexpr bool SEP_ROOM = NT == 1 || D16MAX <= 64;
	const bool sep_p = SEP_ROOM && p.sep_p && !(NT == 1 && D16MAX <= 64);       // (the dual form has no room for separate P columns)
	constexpr uint32_t O_BASE = NT * 128;
	constexpr uint32_t O_STRIDE = NT == 2 ? (SEP_ROOM ? 64 : 128) : 256;
	const uint32_t P_BASE = sep_p ? 384 : 0, P_STRIDE = sep_p ? 64 : 128;

	if (threadIdx.x == 0) {
		tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
		for (int s = 0; s < p.stages; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1); }
		for (int t = 0; t < 2; ++t) { mbar_init(&q_full[t], 1); mbar_init(&s_full[t], 1); mbar_init(&p_full[t], 128); mbar_init(&pv_full[t], 1); mbar_init(&s_free[t], 4); }
		fence_barrier_init();
	}
	constexpr int W_TMA = NT * 4, W_MMA = NT * 4 + 1;     // the issuing warps have the HIGHEST warp ids: the SM's arbiter favours them
	constexpr uint32_t TMEM_COLS_USED = (NT == 1 && D16MAX <= 64) ? 256 : A_TMEM_COLS;
	if (warp == W_MMA) tmem_alloc(tmem_slot, TMEM_COLS_USED);
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem_base = uniform_u32(*tmem_slot);

	if (warp == W_TMA) {
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
	} else if (warp == W_MMA) {
		// ===== MMA issuer: the whole warp runs the (warp-uniform) control flow and the barrier waits, one elected
		// lane issues the tcgen05 instructions, so descriptors and addresses stay in uniform registers.
		const uint32_t idesc_qk = make_idesc_f16(AQ, AK, 0, 0);
		const uint32_t idesc_pv = make_idesc_f16(AQ, p.d16, 0, 1);     // A = P (TMEM), B = V MN-major ([key][d], d contiguous)
		// Descriptors are built once; per MMA only a constant is added (16 columns = 32 B = +2 in the 16-byte
		// address field; the next 64-column chunk is CHUNK_BYTES further).
		const uint64_t qdesc0 = make_smem_desc_sw128(smem_u32(sQ), 16, 1024);
		const uint64_t kdesc0 = make_smem_desc_sw128(smem_u32(sK), 16, 1024);
		const uint64_t vdesc0 = make_smem_desc_sw128(smem_u32(sV), CHUNK_BYTES, 1024);
		const uint32_t tile16 = (uint32_t)tile_bytes >> 4;
		const int nk16 = p.d16 >> 4;
		auto issue_qk = [&](int t, int j) {
			const int s = j % p.stages;
			if (t == 0) mbar_wait(&k_full[s], (uint32_t)(j / p.stages) & 1);
			tc_fence_after();
			if (elect_one()) {
				const uint64_t ad = qdesc0 + (uint64_t)(t * tile16), bd = kdesc0 + (uint64_t)(s * tile16);
				const uint32_t td = tmem_base + t * 128;
				#pragma unroll
				for (int kk = 0; kk < D16MAX / 16; ++kk) {      // unrolled: every offset is an immediate
					const uint32_t off = (uint32_t)((kk >> 2) * (CHUNK_BYTES >> 4) + (kk & 3) * 2);
					if (kk < nk16) umma_f16(td, ad + off, bd + off, idesc_qk, kk ? 1u : 0u);
				}
				umma_commit(&s_full[t]);
				if (t == ntile - 1) umma_commit(&k_empty[s]);
				ATTN_TR(2, j, t * 4 + 3);
			}
			__syncwarp();
		};
		// O_t += P_t V, accumulated IN TENSOR MEMORY over all key blocks (the softmax warps rescale it in
		// place on the rare blocks where the running maximum moves by more than 2^8)
		auto issue_pv = [&](int t, int j) {
			const int s = j % p.stages;
			if (t == 0) mbar_wait(&v_full[s], (uint32_t)(j / p.stages) & 1);
			mbar_wait(&p_full[t], (uint32_t)j & 1);
			tc_fence_after();
			if (elect_one()) {
				ATTN_TR(2, j, t * 4 + 1);
				// V tile: rows = keys (128 B each), 8-row swizzle atoms 1024 B apart (SBO), next 64 d-columns CHUNK_BYTES
				// away (LBO); 16 keys further = +16*128 B. P in tensor memory: 16 keys = 8 packed 32-bit columns.
				const uint64_t bd = vdesc0 + (uint64_t)(s * tile16);
				const uint32_t td = tmem_base + O_BASE + t * O_STRIDE, ta = tmem_base + P_BASE + t * P_STRIDE;
				const int nkk = (min(AK, p.nk - j * AK) + 15) >> 4;
				#pragma unroll
				for (int kk = 0; kk < AK / 16; ++kk)
					if (kk < nkk) umma_f16_ts(td, ta + kk * 8, bd + (uint64_t)(kk * 128), idesc_pv, (j | kk) ? 1u : 0u);
				umma_commit(&pv_full[t]);
				if (t == ntile - 1) umma_commit(&v_empty[s]);
				ATTN_TR(2, j, t * 4 + 2);
			}
			__syncwarp();
		};
		for (int t = 0; t < ntile; ++t) { mbar_wait(&q_full[t], 0); }
		for (int t = 0; t < ntile; ++t) issue_qk(t, 0);
		for (int j = 0; j < p.nblk; ++j) {
			for (int t = 0; t < ntile; ++t) {
				if (sep_p) {
					// S_t is free once the softmax warps hold block j's scores in registers: the next QK^T goes first
					if (j + 1 < p.nblk) { mbar_wait(&s_free[t], (uint32_t)j & 1); issue_qk(t, j + 1); }
					issue_pv(t, j);
				} else {
					issue_pv(t, j);                             // in-order after it: S_t/P_t may be overwritten
					if (j + 1 < p.nblk) issue_qk(t, j + 1);
				}
			}
		}
	} else {
		// ===== softmax: group 0 = warps 0..3 (tile a), group 1 = warps 4..7 (tile b); one query row per thread =====
		const int t = warp >> 2;
		if (t < ntile) {
			const int quarter = warp & 3;
			const int r = quarter * 32 + lane;
			const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
			const uint32_t ts = tmem_base + t * 128 + lane_off;                     // scores of this row
			const uint32_t tp = tmem_base + P_BASE + t * P_STRIDE + lane_off;        // probabilities of this row (packed f16 pairs)
			const uint32_t to = tmem_base + O_BASE + t * O_STRIDE + lane_off;       // output accumulator of this row
			const float sl2 = p.scale_log2;
			float m = -INFINITY, l = 0.f;     // running (possibly stale) maximum in the exp2 domain, running sum
			const bool pingpong = NT == 2 && ntile == 2 && p.pingpong;
			const bool tr = quarter == 0 && lane == 0;     // first warp of the group
			if (pingpong && t == 1) named_bar_arrive(2, 256);      // tile a takes the first turn

			// One key block: two passes over the 128 scores of this row (max, then probabilities), 32 columns
			// at a time, the next tcgen05.ld always in flight while the current chunk is processed. FULL is a
			// compile-time tag so that the common full block carries no per-element masking (ISETP/FSEL).
			auto block = [&](int j, auto full_tag) {
				constexpr bool FULL = decltype(full_tag)::value;
				const int valid = p.nk - j * AK;
				if (tr) ATTN_TR(t, j, 0);
				mbar_wait(&s_full[t], (uint32_t)j & 1);
				tc_fence_after();
				if (tr) ATTN_TR(t, j, 1);
				uint32_t va[32], vb[32];
				float mx4[4] = { -INFINITY, -INFINITY, -INFINITY, -INFINITY };
				float rs4[4] = { 0.f, 0.f, 0.f, 0.f };
				auto maxupd = [&](const uint32_t* v, int c) {
					if (FULL) {
						#pragma unroll
						for (int i = 0; i < 32; i += 8) {      // 3-input max (FMNMX3): half the ALU-pipe work
							mx4[0] = max3f(mx4[0], __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
							mx4[1] = max3f(mx4[1], __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
							mx4[2] = max3f(mx4[2], __uint_as_float(v[i + 4]), __uint_as_float(v[i + 5]));
							mx4[3] = max3f(mx4[3], __uint_as_float(v[i + 6]), __uint_as_float(v[i + 7]));
						}
					} else {
						#pragma unroll
						for (int i = 0; i < 32; ++i) if (c * 32 + i < valid) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(v[i]));
					}
				};
				tmem_ld32(ts, va);
				tmem_ld_wait(); tmem_ld32(ts + 32, vb); maxupd(va, 0);
				tmem_ld_wait(); tmem_ld32(ts + 64, va); maxupd(vb, 1);
				tmem_ld_wait(); tmem_ld32(ts + 96, vb); maxupd(va, 2);
				tmem_ld_wait(); tmem_ld32(ts, va);      maxupd(vb, 3);
				const float m_blk = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * sl2;
				// Lazy rescaling: keep the old maximum while the block maximum exceeds it by < 2^8 (p <= 256 is
				// exact enough in f16 and cannot overflow the f32 sums); otherwise move it and rescale O and l.
				if (j == 0) m = m_blk;
				else {
					const bool need = m_blk > m + 8.0f;
					if (__any_sync(0xffffffffu, need)) {
						const float m_new = need ? m_blk : m;
						const float corr = ex2_approx(m - m_new);
						m = m_new;
						l *= corr;
						mbar_wait(&pv_full[t], (uint32_t)(j - 1) & 1);     // all PV products so far have landed in O
						tc_fence_after();
						#pragma unroll
						for (int c0 = 0; c0 < D16MAX; c0 += 16) {
							if (c0 < p.d16) {
								uint32_t o[16];
								tmem_ld16(to + c0, o);
								tmem_ld_wait();
								#pragma unroll
								for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * corr);
								tmem_st16(to + c0, o);
							}
						}
					}
				}
				// Ping-pong: the exp pass (MUFU-bound) of the two tiles strictly alternates, so one tile's
				// QK/PV products and max pass run under the other tile's exponentials instead of both tiles
				// contending for the MUFU in lockstep and then idling together while the tensor core works.
				if (tr) ATTN_TR(t, j, 2);
				if (pingpong) named_bar_sync(2 + t, 256);
				if (tr) ATTN_TR(t, j, 3);
				// separate P columns: the previous block's PV product must have consumed P_t before it is rewritten
				if (sep_p && j > 0) { mbar_wait(&pv_full[t], (uint32_t)(j - 1) & 1); tc_fence_after(); }
				const float mneg = -m;
				// Exp pass, software-pipelined over the four 32-column chunks with three register buffers: the
				// exponentials (FFMA + MUFU.EX2, in place) of chunk c+1 are issued BEFORE the sum / f16 pack /
				// tcgen05.st of chunk c, so the MUFU stream never drains at a chunk boundary.
				uint32_t vc[32];
				auto exps = [&](uint32_t* v, int c) {
					#pragma unroll
					for (int i = 0; i < 32; ++i) {
						const float xs = fmaf(__uint_as_float(v[i]), sl2, mneg);
						float e = (i & 7) < N_POLY ? ex2_poly(xs) : ex2_approx(xs);
						if (!FULL) { if (c * 32 + i >= valid) e = 0.f; }
						v[i] = __float_as_uint(e);
					}
				};
				auto finish = [&](const uint32_t* v, int c) {
					uint32_t packed[16];
					#pragma unroll
					for (int i = 0; i < 32; i += 2) {
						const float p0 = __uint_as_float(v[i]), p1 = __uint_as_float(v[i + 1]);
						rs4[(i >> 1) & 3] += p0 + p1;
						__half2 hh = __floats2half2_rn(p0, p1);
						packed[i >> 1] = *reinterpret_cast<uint32_t*>(&hh);
					}
					tmem_st16(tp + c * 16, packed);    // (aliased layout: columns [16c, 16c+16) lie inside scores that were already read)
				};
				tmem_ld_wait(); tmem_ld32(ts + 32, vb); exps(va, 0);
				tmem_ld_wait(); tmem_ld32(ts + 64, vc); exps(vb, 1); finish(va, 0);
				tmem_ld_wait(); tmem_ld32(ts + 96, va); exps(vc, 2); finish(vb, 1);
				tmem_ld_wait();
				if (sep_p) {                            // every score of the block is in registers: S_t may be overwritten
					tc_fence_before();
					__syncwarp();
					if (lane == 0) mbar_arrive(&s_free[t]);
				}
				exps(va, 3); finish(vc, 2);
				finish(va, 3);
				if (tr) ATTN_TR(t, j, 4);
				if (pingpong && (t == 0 || j + 1 < p.nblk)) named_bar_arrive(2 + (t ^ 1), 256);
				l += (rs4[0] + rs4[1]) + (rs4[2] + rs4[3]);
				tmem_st_wait();
				tc_fence_before();
				mbar_arrive(&p_full[t]);
				if (tr) ATTN_TR(t, j, 5);
			};
			const int nfull = p.nk / AK;      // a partial block can only be the last one
			for (int j = 0; j < nfull; ++j) block(j, std::true_type{});
			if (nfull < p.nblk) block(nfull, std::false_type{});

			// epilogue: O / l, written token-major ([token][head][d]) for the output projection GEMM
			mbar_wait(&pv_full[t], (uint32_t)(p.nblk - 1) & 1);
			tc_fence_after();
			const long long tok = (long long)q0 + t * AQ + r;
			const float inv = l > 0.f ? 1.0f / l : 0.f;
			__half* op = (__half*)p.o + tok * p.so_t + (long long)h * p.so_h + (long long)b * p.so_b;
			const bool vec = ((((uintptr_t)op) & 15) == 0);
			#pragma unroll
			for (int c0 = 0; c0 < D16MAX; c0 += 16) {
				if (c0 < p.d16) {
					uint32_t o[16];
					tmem_ld16(to + c0, o);
					tmem_ld_wait();
					if (tok < p.nq) {
						#pragma unroll
						for (int h8 = 0; h8 < 16; h8 += 8) {
							const int cc = c0 + h8;
							if (cc < p.d) {
								if (vec && cc + 8 <= p.d) {
									uint4 o4; __half2* hp = reinterpret_cast<__half2*>(&o4);
									#pragma unroll
									for (int i = 0; i < 4; ++i) hp[i] = __floats2half2_rn(__uint_as_float(o[h8 + 2 * i]) * inv, __uint_as_float(o[h8 + 2 * i + 1]) * inv);
									*reinterpret_cast<uint4*>(op + cc) = o4;
								} else {
									#pragma unroll
									for (int i = 0; i < 8; ++i) if (cc + i < p.d) op[cc + i] = __float2half_rn(__uint_as_float(o[h8 + i]) * inv);
								}
							}
						}
					}
				}
			}
		}
	}
	tc_fence_before();
	__syncthreads();
	if (warp == W_MMA) { tc_fence_after(); tmem_dealloc(tmem_base, TMEM_COLS_USED); }
}



// ------------------------------------------------------------------ heads up to 64 wide: key-split form
// For d <= 64 the exponentials bound the kernel (MUFU.EX2: 16 per clock and SM, i.e. 1024 clk per 128 x 128 score tile against
// 2 * (d16/16) * 64 clk of MMA). The one-row-per-thread forms above keep ONE (ping-pong) or TWO (dual) softmax warps per SM
// sub-partition, and each of them spends half of its time outside the exponential pass (waiting for the products, loading
// scores twice, the maximum pass): the MUFU pipe idles 48 % of the time (ncu, profiles/r1_ncu_attention.md).
// Here a CTA still owns one tile of 128 query rows, but EIGHT softmax warps work on every key block: warps 0..3 take the
// keys [0, 64) of the block, warps 4..7 the keys [64, 128). The two halves never talk to each other inside the loop:
//   * each half keeps its own running maximum and sum per row and its own output accumulator in tensor memory
//     (O_a, O_b: the PV product of a block is issued as two K = 64 products), like a split-KV decode;
//   * a thread reads its 64 scores ONCE (two tcgen05.ld), keeps them in registers for the maximum and the exponentials,
//     and writes the f16 probabilities over the first half of its own score columns;
//   * the halves are merged in the epilogue through 2 KB of shared memory: O = (O_a w_a + O_b w_b) / (l_a w_a + l_b w_b).
// Two such CTAs share an SM (256 tensor-memory columns, <= 113 KB of shared memory each): four softmax warps per SM
// sub-partition keep the MUFU pipe fed while the others wait, load or pack.
// STAG: the two halves run in ANTI-PHASE. Each half has its own score columns' barrier (QK^T is issued as two N = 64 products),
// its own probabilities / PV barriers, and the issuing warp serves them alternately: PV_a(j), QK_a(j+1) when half a arrives,
// PV_b(j), QK_b(j+1) when half b arrives; half b's first product is only issued when half a has finished its first block.
// While one half waits for the tensor pipe the other one is in its exponential pass, inside one CTA.
// PK: 0 = scalar softmax arithmetic (round-2 form, N_POLY of every 8 exponentials on the FMA pipe);
//     1 = packed f32x2 arithmetic (FFMA2 scale-and-subtract, FADD2 row sums, packed polynomial: N_POLY PAIRS of every 8 pairs);
//     2 = packed, and the row sums come from the TENSOR CORE: every PV product is followed by a product of the same probabilities
//         with a tile of ones (N = 16) into 16 tensor-memory columns right of the half's accumulator, so the sum is taken over the
//         f16-rounded probabilities that the PV product really used, and the softmax loop holds no additions at all. Needs
//         d16 + 16 <= 64 columns per half (heads up to 48 wide: SD1.x level 0).
template <int N_POLY, bool STAG, int PK = 0>
__global__ void __launch_bounds__(320, 2)
attn_split_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
	const AttnParams p)
{
	extern __shared__ __align__(1024) uint8_t smem_raw[];
	uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
	const int tile_bytes = p.dchunks * CHUNK_BYTES;       // dchunks == 1 here (d <= 64)
	uint8_t* sQ = smem;
	uint8_t* sK = sQ + tile_bytes;
	uint8_t* sV = sK + (size_t)p.stages * tile_bytes;
	uint8_t* sOnes = sV + (size_t)p.stages * tile_bytes;             // PK == 2: 16 key rows of 128 B, every f16 = 1.0 (any swizzle of it is itself)
	uint64_t* bars = (uint64_t*)(sOnes + (PK == 2 ? 2048 : 0));
	uint64_t* q_full = bars;                        // [1]
	uint64_t* k_full = q_full + 1;                  // [stages]
	uint64_t* k_empty = k_full + A_MAX_STAGES;
	uint64_t* v_full = k_empty + A_MAX_STAGES;
	uint64_t* v_empty = v_full + A_MAX_STAGES;
	uint64_t* s_full = v_empty + A_MAX_STAGES;      // [2]   QK done (STAG: per half)
	uint64_t* p_full = s_full + 2;                  // [2]   probabilities written (both halves on [0], 256 arrivals; STAG: per half, 128)
	uint64_t* pv_full = p_full + 2;                 // [2]   PV products of a block done (STAG: per half)
	uint32_t* tmem_slot = (uint32_t*)(pv_full + 2);
	float2* exch = (float2*)(tmem_slot + 2);        // [2 halves][128 rows] (running maximum, running sum)

	const int warp = warp_id_uniform(), lane = threadIdx.x & 31;
	const int q0 = blockIdx.x * AQ, h = blockIdx.y, b = blockIdx.z;
	// clock probe of one mid-run CTA (tools/attn_trace.cu): SM cycles and nanoseconds it lived, i.e. the SM clock under this kernel
	const bool probe = p.trace && threadIdx.x == 0 && blockIdx.x == gridDim.x / 2 && blockIdx.y == gridDim.y / 2 && blockIdx.z == gridDim.z / 2;
	long long probe_c0 = 0; unsigned long long probe_g0 = 0;
	if (probe) { probe_c0 = clock64(); asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(probe_g0)); }
	constexpr uint32_t O_BASE = 128, O_STRIDE = 64;   // TMEM: S at 0 (P_a over columns [0,32), P_b over [64,96)), O_a at 128, O_b at 192
	constexpr int W_TMA = 8, W_MMA = 9;

	if (threadIdx.x == 0) {
		tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
		for (int s = 0; s < p.stages; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1); }
		mbar_init(q_full, 1);
		for (int t = 0; t < 2; ++t) { mbar_init(&s_full[t], 1); mbar_init(&p_full[t], STAG ? 128 : 256); mbar_init(&pv_full[t], 1); }
		fence_barrier_init();
	}
	if (PK == 2) {
		if (threadIdx.x < 128) reinterpret_cast<uint4*>(sOnes)[threadIdx.x] = make_uint4(0x3C003C00u, 0x3C003C00u, 0x3C003C00u, 0x3C003C00u);
		fence_proxy_async();
	}
	if (warp == W_MMA) tmem_alloc(tmem_slot, 256);
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem_base = uniform_u32(*tmem_slot);

	if (warp == W_TMA) {
		if (lane == 0) {
			mbar_expect_tx(q_full, tile_bytes);
			tma_load_4d(sQ, &tmQ, q_full, 0, q0, h, b);
			for (int j = 0; j < p.nblk; ++j) {
				const int s = j % p.stages; const uint32_t ph = (uint32_t)(j / p.stages) & 1;
				mbar_wait_parked(&k_empty[s], ph ^ 1);
				mbar_expect_tx(&k_full[s], tile_bytes);
				tma_load_4d(sK + (size_t)s * tile_bytes, &tmK, &k_full[s], 0, j * AK, h, b);
				mbar_wait_parked(&v_empty[s], ph ^ 1);
				mbar_expect_tx(&v_full[s], tile_bytes);
				tma_load_4d(sV + (size_t)s * tile_bytes, &tmV, &v_full[s], 0, j * AK, h, b);
			}
		}
	} else if (warp == W_MMA) {
		const uint32_t idesc_qk = make_idesc_f16(AQ, AK, 0, 0);
		const uint32_t idesc_pv = make_idesc_f16(AQ, p.d16, 0, 1);     // A = P (TMEM), B = V MN-major ([key][d], d contiguous)
		const uint64_t qdesc0 = make_smem_desc_sw128(smem_u32(sQ), 16, 1024);
		const uint64_t kdesc0 = make_smem_desc_sw128(smem_u32(sK), 16, 1024);
		const uint64_t vdesc0 = make_smem_desc_sw128(smem_u32(sV), CHUNK_BYTES, 1024);
		const uint32_t tile16 = (uint32_t)tile_bytes >> 4;
		const int nk16 = p.d16 >> 4;
		const uint32_t idesc_sum = make_idesc_f16(AQ, 16, 0, 1);                     // PK == 2: row sums = P x ones
		const uint64_t odesc = make_smem_desc_sw128(smem_u32(sOnes), CHUNK_BYTES, 1024);
		if (STAG) {
			const uint32_t idesc_qk64 = make_idesc_f16(AQ, 64, 0, 0);
			// S_hf = Q K(j)[64 hf .. 64 hf + 64)^T : the K tile's rows 64 hf.. are 64 * 128 B further (+512 in 16-byte units)
			auto qk_half = [&](int j, int hf) {
				const uint64_t bd = kdesc0 + (uint64_t)((j % p.stages) * tile16) + (uint64_t)(hf * 512);
				#pragma unroll
				for (int kk = 0; kk < 4; ++kk)
					if (kk < nk16) umma_f16(tmem_base + hf * 64, qdesc0 + (uint64_t)(kk * 2), bd + (uint64_t)(kk * 2), idesc_qk64, kk ? 1u : 0u);
				umma_commit(&s_full[hf]);
			};
			auto pv_half = [&](int j, int hf) {
				const int s = j % p.stages;
				const int valid = min(AK, p.nk - j * AK);
				const uint64_t bd = vdesc0 + (uint64_t)(s * tile16) + (uint64_t)(hf * 512);
				const uint32_t td = tmem_base + O_BASE + hf * O_STRIDE, ta = tmem_base + hf * 64;
				const int nkk = (min(64, max(0, valid - hf * 64)) + 15) >> 4;
				#pragma unroll
				for (int kk = 0; kk < 4; ++kk)
					if (kk < nkk) umma_f16_ts(td, ta + kk * 8, bd + (uint64_t)(kk * 128), idesc_pv, (j | kk) ? 1u : 0u);
				umma_commit(&pv_full[hf]);
			};
			mbar_wait(q_full, 0);
			mbar_wait(&k_full[0], 0);
			tc_fence_after();
			if (elect_one()) qk_half(0, 0);
			__syncwarp();
			for (int j = 0; j < p.nblk; ++j) {
				const int s = j % p.stages;
				const bool more = j + 1 < p.nblk;
				mbar_wait(&v_full[s], (uint32_t)(j / p.stages) & 1);
				if (more) mbar_wait(&k_full[(j + 1) % p.stages], (uint32_t)((j + 1) / p.stages) & 1);
				mbar_wait(&p_full[0], (uint32_t)j & 1);
				tc_fence_after();
				if (elect_one()) {
					pv_half(j, 0);
					if (j == 0) qk_half(0, 1);               // half b starts one softmax pass behind half a
					if (more) qk_half(j + 1, 0);
				}
				__syncwarp();
				mbar_wait(&p_full[1], (uint32_t)j & 1);
				tc_fence_after();
				if (elect_one()) {
					pv_half(j, 1);
					umma_commit(&v_empty[s]);
					umma_commit(&k_empty[s]);                // QK_a(j) and QK_b(j) were issued before this point
					if (more) qk_half(j + 1, 1);
				}
				__syncwarp();
			}
		} else {
		// The issuing warp is the critical path between the softmax of block j and that of block j + 1 (PV(j), then QK(j+1)
		// into the same score columns). Everything that can be waited for early is waited for BEFORE the probabilities
		// arrive (V(j) and K(j+1) resident); after p_full only one fence and one elected section remain.
		mbar_wait(q_full, 0);
		mbar_wait(&k_full[0], 0);
		tc_fence_after();
		if (elect_one()) {
			#pragma unroll
			for (int kk = 0; kk < 4; ++kk)
				if (kk < nk16) umma_f16(tmem_base, qdesc0 + (uint64_t)(kk * 2), kdesc0 + (uint64_t)(kk * 2), idesc_qk, kk ? 1u : 0u);
			umma_commit(&s_full[0]);
			umma_commit(&k_empty[0]);
			ATTN_TR(2, 0, 3);
		}
		__syncwarp();
		for (int j = 0; j < p.nblk; ++j) {
			const int s = j % p.stages, s1 = (j + 1) % p.stages;
			const int valid = min(AK, p.nk - j * AK);
			const bool more = j + 1 < p.nblk;
			mbar_wait(&v_full[s], (uint32_t)(j / p.stages) & 1);
			if (more) mbar_wait(&k_full[s1], (uint32_t)((j + 1) / p.stages) & 1);
			if (lane == 0) ATTN_TR(3, j, 4);
			mbar_wait(&p_full[0], (uint32_t)j & 1);
			tc_fence_after();
			if (elect_one()) {
				ATTN_TR(2, j, 1);
				// V tile: rows = keys (128 B each), 16 keys further = +16 * 128 B; the second half starts 64 keys in.
				// P in tensor memory: 16 keys = 8 packed 32-bit columns, half b's probabilities start at column 64.
				#pragma unroll
				for (int hf = 0; hf < 2; ++hf) {
					const uint64_t bd = vdesc0 + (uint64_t)(s * tile16) + (uint64_t)(hf * 512);
					const uint32_t td = tmem_base + O_BASE + hf * O_STRIDE, ta = tmem_base + hf * 64;
					const int nkk = (min(64, max(0, valid - hf * 64)) + 15) >> 4;
					#pragma unroll
					for (int kk = 0; kk < 4; ++kk)
						if (kk < nkk) umma_f16_ts(td, ta + kk * 8, bd + (uint64_t)(kk * 128), idesc_pv, (j | kk) ? 1u : 0u);
					if (PK == 2) {
						#pragma unroll
						for (int kk = 0; kk < 4; ++kk)
							if (kk < nkk) umma_f16_ts(td + p.d16, ta + kk * 8, odesc, idesc_sum, (j | kk) ? 1u : 0u);
					}
				}
				umma_commit(&pv_full[0]);
				umma_commit(&v_empty[s]);
				ATTN_TR(2, j, 2);
				if (more) {                              // in order after both products: S / P may be overwritten
					const uint64_t bd = kdesc0 + (uint64_t)(s1 * tile16);
					#pragma unroll
					for (int kk = 0; kk < 4; ++kk)
						if (kk < nk16) umma_f16(tmem_base, qdesc0 + (uint64_t)(kk * 2), bd + (uint64_t)(kk * 2), idesc_qk, kk ? 1u : 0u);
					umma_commit(&s_full[0]);
					umma_commit(&k_empty[s1]);
					ATTN_TR(2, j + 1, 3);
				}
			}
			__syncwarp();
		}
		}
	} else {
		// ===== softmax: warp = quarter (TMEM lanes) + 4 * half (key columns); one (row, half) per thread =====
		const int hf = warp >> 2, quarter = warp & 3;
		const int r = quarter * 32 + lane;
		const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
		const uint32_t ts = tmem_base + lane_off + hf * 64;                      // my 64 scores; the probabilities go over their first 32 columns
		const uint32_t to = tmem_base + O_BASE + hf * O_STRIDE + lane_off;       // my half's output accumulator of this row
		const float sl2 = p.scale_log2;
		float m = -INFINITY, l = 0.f;
		bool first = true;
		const bool tr = PK == 0 && quarter == 0 && lane == 0;     // timeline probe (tools/attn_trace.cu): scalar form only
		const int hb = STAG ? hf : 0;                     // barrier set of this thread

		auto block = [&](int j, auto full_tag) {
			constexpr bool FULL = decltype(full_tag)::value;
			const int valid = FULL ? 64 : min(64, max(0, p.nk - j * AK - hf * 64));      // valid keys of my half
			if (tr) ATTN_TR(hf, j, 0);
			mbar_wait_parked(&s_full[hb], (uint32_t)j & 1);
			tc_fence_after();
			if (tr) ATTN_TR(hf, j, 1);
			uint32_t va[32], vb[32];
			if (!FULL && valid <= 0) {                                           // nothing of this block belongs to my half
				#pragma unroll
				for (int i = 0; i < 16; ++i) va[i] = 0u;
				tmem_st16(ts, va); tmem_st16(ts + 16, va);
				tmem_st_wait(); tc_fence_before(); mbar_arrive(&p_full[hb]);
				return;
			}
			tmem_ld32(ts, va); tmem_ld32(ts + 32, vb);
			tmem_ld_wait();
			float mx4[4] = { -INFINITY, -INFINITY, -INFINITY, -INFINITY };
			if (FULL) {
				#pragma unroll
				for (int i = 0; i < 32; i += 8) {
					mx4[0] = max3f(mx4[0], __uint_as_float(va[i]), __uint_as_float(va[i + 1]));
					mx4[1] = max3f(mx4[1], __uint_as_float(va[i + 2]), __uint_as_float(va[i + 3]));
					mx4[2] = max3f(mx4[2], __uint_as_float(va[i + 4]), __uint_as_float(va[i + 5]));
					mx4[3] = max3f(mx4[3], __uint_as_float(va[i + 6]), __uint_as_float(va[i + 7]));
					mx4[0] = max3f(mx4[0], __uint_as_float(vb[i]), __uint_as_float(vb[i + 1]));
					mx4[1] = max3f(mx4[1], __uint_as_float(vb[i + 2]), __uint_as_float(vb[i + 3]));
					mx4[2] = max3f(mx4[2], __uint_as_float(vb[i + 4]), __uint_as_float(vb[i + 5]));
					mx4[3] = max3f(mx4[3], __uint_as_float(vb[i + 6]), __uint_as_float(vb[i + 7]));
				}
			} else {
				#pragma unroll
				for (int i = 0; i < 32; ++i) {
					if (i < valid) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(va[i]));
					if (32 + i < valid) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(vb[i]));
				}
			}
			const float m_blk = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * sl2;
			if (tr) ATTN_TR(hf, j, 2);
			// Lazy rescaling (as above): keep the old maximum while the block maximum exceeds it by < 2^8.
			if (first) { m = m_blk; first = false; }
			else {
				const bool need = m_blk > m + 8.0f;
				if (__any_sync(0xffffffffu, need)) {
					const float m_new = need ? m_blk : m;
					const float corr = ex2_approx(m - m_new);
					m = m_new;
					l *= corr;
					mbar_wait(&pv_full[hb], (uint32_t)(j - 1) & 1);     // the PV products so far have landed in O
					tc_fence_after();
					#pragma unroll
					for (int c0 = 0; c0 < 64; c0 += 16) {
						if (c0 < p.d16 + (PK == 2 ? 16 : 0)) {                   // PK == 2: the sum columns follow the accumulator
							uint32_t o[16];
							tmem_ld16(to + c0, o);
							tmem_ld_wait();
							#pragma unroll
							for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * corr);
							tmem_st16(to + c0, o);
						}
					}
				}
			}
			const float mneg = -m;
			if (tr) ATTN_TR(hf, j, 3);
			if constexpr (PK > 0) {
				const uint64_t SL = pk2(sl2, sl2), MN = pk2(mneg, mneg);
				uint64_t acc0 = 0ull, acc1 = 0ull;
				auto pass = [&](uint32_t* v, int c) {
					uint32_t packed[16];
					#pragma unroll
					for (int i = 0; i < 32; i += 2) {
						const int pi = i >> 1;
						const uint64_t x2 = fma2(pk2u(v[i], v[i + 1]), SL, MN);
						float e0, e1;
						if (N_POLY > 0 && (((pi & 7) * N_POLY) & 7) < N_POLY) ex2_poly2(x2, e0, e1);
						else { float x0, x1; unpk2(x2, x0, x1); e0 = ex2_approx(x0); e1 = ex2_approx(x1); }
						if (!FULL) { if (c * 32 + i >= valid) e0 = 0.f; if (c * 32 + i + 1 >= valid) e1 = 0.f; }
						if (PK == 1) { if (pi & 1) acc1 = add2(acc1, pk2(e0, e1)); else acc0 = add2(acc0, pk2(e0, e1)); }
						__half2 hh = __floats2half2_rn(e0, e1);
						packed[pi] = *reinterpret_cast<uint32_t*>(&hh);
					}
					tmem_st16(ts + c * 16, packed);
				};
				pass(va, 0); pass(vb, 1);
				if (PK == 1) { float a0, a1; unpk2(add2(acc0, acc1), a0, a1); l += a0 + a1; }
			} else {
				float rs4[4] = { 0.f, 0.f, 0.f, 0.f };
				auto exps = [&](uint32_t* v, int c) {
					#pragma unroll
					for (int i = 0; i < 32; ++i) {
						const float xs = fmaf(__uint_as_float(v[i]), sl2, mneg);
						float e = (i & 7) < N_POLY ? ex2_poly(xs) : ex2_approx(xs);
						if (!FULL) { if (c * 32 + i >= valid) e = 0.f; }
						v[i] = __float_as_uint(e);
					}
				};
				auto finish = [&](const uint32_t* v, int c) {
					uint32_t packed[16];
					#pragma unroll
					for (int i = 0; i < 32; i += 2) {
						const float p0 = __uint_as_float(v[i]), p1 = __uint_as_float(v[i + 1]);
						rs4[(i >> 1) & 3] += p0 + p1;
						__half2 hh = __floats2half2_rn(p0, p1);
						packed[i >> 1] = *reinterpret_cast<uint32_t*>(&hh);
					}
					tmem_st16(ts + c * 16, packed);
				};
				exps(va, 0); exps(vb, 1);
				finish(va, 0); finish(vb, 1);
				l += (rs4[0] + rs4[1]) + (rs4[2] + rs4[3]);
			}
			if (tr) ATTN_TR(hf, j, 4);
			tmem_st_wait();
			tc_fence_before();
			mbar_arrive(&p_full[hb]);
			if (tr) ATTN_TR(hf, j, 5);
		};
		const int nfull = p.nk / AK;      // a partial block can only be the last one
		for (int j = 0; j < nfull; ++j) block(j, std::true_type{});
		if (nfull < p.nblk) block(nfull, std::false_type{});

		// epilogue: merge the halves. Every thread publishes (m, l) of its (row, half), then takes the 16-column chunks
		// c with (c & 1) == half of BOTH accumulators of its row.
		mbar_wait(&pv_full[0], (uint32_t)(p.nblk - 1) & 1);
		if (STAG) mbar_wait(&pv_full[1], (uint32_t)(p.nblk - 1) & 1);
		tc_fence_after();
		if (PK == 2) {                                   // the row sum of my half: any of the 16 columns right of the accumulator
			uint32_t t16[16];
			tmem_ld16(to + p.d16, t16);
			tmem_ld_wait();
			l = __uint_as_float(t16[0]);
		}
		exch[hf * 128 + r] = make_float2(m, l);
		named_bar_sync(1, 256);
		const float2 oth = exch[(hf ^ 1) * 128 + r];
		const float M = fmaxf(m, oth.x);
		const float w_me = first ? 0.f : ex2_approx(m - M), w_ot = (oth.x == -INFINITY) ? 0.f : ex2_approx(oth.x - M);
		const float L = l * w_me + oth.y * w_ot;
		const float inv = L > 0.f ? 1.0f / L : 0.f;
		const float s_me = w_me * inv, s_ot = w_ot * inv;
		const uint32_t to_ot = tmem_base + O_BASE + (hf ^ 1) * O_STRIDE + lane_off;
		const long long tok = (long long)q0 + r;
		__half* op = (__half*)p.o + tok * p.so_t + (long long)h * p.so_h + (long long)b * p.so_b;
		const bool vec = ((((uintptr_t)op) & 15) == 0);
		#pragma unroll
		for (int c = 0; c < 4; ++c) {
			const int c0 = c * 16;
			if ((c & 1) == hf && c0 < p.d16) {
				uint32_t oa[16], ob[16];
				tmem_ld16(to + c0, oa); tmem_ld16(to_ot + c0, ob);
				tmem_ld_wait();
				if (tok < p.nq) {
					float o[16];
					#pragma unroll
					for (int i = 0; i < 16; ++i) o[i] = __uint_as_float(oa[i]) * s_me + __uint_as_float(ob[i]) * s_ot;
					#pragma unroll
					for (int h8 = 0; h8 < 16; h8 += 8) {
						const int cc = c0 + h8;
						if (cc < p.d) {
							if (vec && cc + 8 <= p.d) {
								uint4 o4; __half2* hp = reinterpret_cast<__half2*>(&o4);
								#pragma unroll
								for (int i = 0; i < 4; ++i) hp[i] = __floats2half2_rn(o[h8 + 2 * i], o[h8 + 2 * i + 1]);
								*reinterpret_cast<uint4*>(op + cc) = o4;
							} else {
								#pragma unroll
								for (int i = 0; i < 8; ++i) if (cc + i < p.d) op[cc + i] = __float2half_rn(o[h8 + i]);
							}
						}
					}
				}
			}
		}
	}
	tc_fence_before();
	__syncthreads();
	if (probe) { unsigned long long g1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1)); p.trace[2046] = clock64() - probe_c0; p.trace[2047] = (long long)(g1 - probe_g0); }
	if (warp == W_MMA) { tc_fence_after(); tmem_dealloc(tmem_base, 256); }
}



// ------------------------------------------------------------------ heads up to 64 wide: quarter-split form, double-buffered scores
// The key-split kernel above still serialises, per CTA, softmax(j) -> PV(j) -> QK(j+1) -> softmax(j+1): its eight softmax warps wait
// ~950 clk per key block for the issuing warp and the tensor pipe (timeline in profiles/r2_attention_microbench.md). Here ONE CTA per
// SM owns all 512 tensor-memory columns:
//   S0 | S1   two score buffers: QK(j+2) is issued as soon as PV(j) has been issued, i.e. the scores of the next TWO blocks
//             are normally complete before the softmax warps ask for them -- the products leave the critical path;
//   O_0..O_3  one output accumulator per key quarter.
// SIXTEEN softmax warps: warp = lane group (TMEM lanes) + 4 * key quarter; a thread owns 32 scores of one row per block (one
// tcgen05.ld), its own running maximum / sum, and writes 16 packed probability columns over the start of its own scores. The
// warps of a sub-partition only meet at the MUFU pipe: four of them per sub-partition keep it fed while others load, pack or wait.
// Epilogue: the four partial (m, l, O) of a row are merged through 4 KB of shared memory.
// NKQ key parts per block (4: sixteen softmax warps, 32 scores per thread; 2: eight softmax warps, 64 scores per thread)
template <int N_POLY, int NKQ>
__global__ void __launch_bounds__(128 * NKQ + 64, 1)
attn_split4_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
	const AttnParams p)
{
	extern __shared__ __align__(1024) uint8_t smem_raw[];
	uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
	const int tile_bytes = p.dchunks * CHUNK_BYTES;       // dchunks == 1 here (d <= 64)
	uint8_t* sQ = smem;
	uint8_t* sK = sQ + tile_bytes;
	uint8_t* sV = sK + (size_t)p.stages * tile_bytes;
	uint64_t* bars = (uint64_t*)(sV + (size_t)p.stages * tile_bytes);
	uint64_t* q_full = bars;                        // [1]
	uint64_t* k_full = q_full + 1;                  // [stages]
	uint64_t* k_empty = k_full + A_MAX_STAGES;
	uint64_t* v_full = k_empty + A_MAX_STAGES;
	uint64_t* v_empty = v_full + A_MAX_STAGES;
	uint64_t* s_full = v_empty + A_MAX_STAGES;      // [2]   QK into score buffer b done
	uint64_t* p_full = s_full + 2;                  // [2]   probabilities of buffer b written (128 NKQ arrivals)
	uint64_t* pv_full = p_full + 2;                 // [1]   PV products of a block done
	uint32_t* tmem_slot = (uint32_t*)(pv_full + 1);
	float2* exch = (float2*)(tmem_slot + 2);        // [4 quarters][128 rows] (running maximum, running sum)

	const int warp = warp_id_uniform(), lane = threadIdx.x & 31;
	const int q0 = blockIdx.x * AQ, h = blockIdx.y, b = blockIdx.z;
	constexpr uint32_t O_BASE = 256, O_STRIDE = 64;
	constexpr int W_TMA = 4 * NKQ, W_MMA = 4 * NKQ + 1;
	constexpr int EPT = 128 / NKQ;                   // scores per thread and block

	if (threadIdx.x == 0) {
		tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
		for (int s = 0; s < p.stages; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1); }
		mbar_init(q_full, 1); mbar_init(pv_full, 1);
		for (int t = 0; t < 2; ++t) { mbar_init(&s_full[t], 1); mbar_init(&p_full[t], 128 * NKQ); }
		fence_barrier_init();
	}
	if (warp == W_MMA) tmem_alloc(tmem_slot, 512);
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem_base = uniform_u32(*tmem_slot);

	if (warp == W_TMA) {
		if (lane == 0) {
			mbar_expect_tx(q_full, tile_bytes);
			tma_load_4d(sQ, &tmQ, q_full, 0, q0, h, b);
			// K runs two blocks ahead of V (QK(j+2) is issued together with PV(j)): issue order K0 K1 [V0 K2] [V1 K3] ...
			auto load_k = [&](int j) {
				const int s = j % p.stages; const uint32_t ph = (uint32_t)(j / p.stages) & 1;
				mbar_wait_parked(&k_empty[s], ph ^ 1);
				mbar_expect_tx(&k_full[s], tile_bytes);
				tma_load_4d(sK + (size_t)s * tile_bytes, &tmK, &k_full[s], 0, j * AK, h, b);
			};
			auto load_v = [&](int j) {
				const int s = j % p.stages; const uint32_t ph = (uint32_t)(j / p.stages) & 1;
				mbar_wait_parked(&v_empty[s], ph ^ 1);
				mbar_expect_tx(&v_full[s], tile_bytes);
				tma_load_4d(sV + (size_t)s * tile_bytes, &tmV, &v_full[s], 0, j * AK, h, b);
			};
			load_k(0);
			if (p.nblk > 1) load_k(1);
			for (int j = 0; j < p.nblk; ++j) { load_v(j); if (j + 2 < p.nblk) load_k(j + 2); }
		}
	} else if (warp == W_MMA) {
		const uint32_t idesc_qk = make_idesc_f16(AQ, AK, 0, 0);
		const uint32_t idesc_pv = make_idesc_f16(AQ, p.d16, 0, 1);     // A = P (TMEM), B = V MN-major ([key][d], d contiguous)
		const uint64_t qdesc0 = make_smem_desc_sw128(smem_u32(sQ), 16, 1024);
		const uint64_t kdesc0 = make_smem_desc_sw128(smem_u32(sK), 16, 1024);
		const uint64_t vdesc0 = make_smem_desc_sw128(smem_u32(sV), CHUNK_BYTES, 1024);
		const uint32_t tile16 = (uint32_t)tile_bytes >> 4;
		const int nk16 = p.d16 >> 4;
		auto qk_mmas = [&](int j) {          // called by the elected lane: S[j & 1] = Q K(j)^T
			const int s = j % p.stages;
			const uint64_t bd = kdesc0 + (uint64_t)(s * tile16);
			const uint32_t td = tmem_base + (uint32_t)(j & 1) * 128;
			#pragma unroll
			for (int kk = 0; kk < 4; ++kk)
				if (kk < nk16) umma_f16(td, qdesc0 + (uint64_t)(kk * 2), bd + (uint64_t)(kk * 2), idesc_qk, kk ? 1u : 0u);
			umma_commit(&s_full[j & 1]);
			umma_commit(&k_empty[s]);
		};
		mbar_wait(q_full, 0);
		mbar_wait(&k_full[0], 0);
		if (p.nblk > 1) mbar_wait(&k_full[1 % p.stages], (uint32_t)(1 / p.stages) & 1);
		tc_fence_after();
		if (elect_one()) { qk_mmas(0); if (p.nblk > 1) qk_mmas(1); }
		__syncwarp();
		for (int j = 0; j < p.nblk; ++j) {
			const int s = j % p.stages;
			const int valid = min(AK, p.nk - j * AK);
			const bool more = j + 2 < p.nblk;
			mbar_wait(&v_full[s], (uint32_t)(j / p.stages) & 1);
			if (more) mbar_wait(&k_full[(j + 2) % p.stages], (uint32_t)((j + 2) / p.stages) & 1);
			mbar_wait(&p_full[j & 1], (uint32_t)(j >> 1) & 1);
			tc_fence_after();
			if (elect_one()) {
				// V tile: rows = keys (128 B each), 16 keys further = +16 * 128 B; quarter q starts 32 keys * q in.
				// P in tensor memory: 16 keys = 8 packed 32-bit columns, quarter q's probabilities start at column 32 q of S[j & 1].
				#pragma unroll
				for (int q = 0; q < NKQ; ++q) {
					const uint64_t bd = vdesc0 + (uint64_t)(s * tile16) + (uint64_t)(q * EPT * 8);
					const uint32_t td = tmem_base + O_BASE + q * O_STRIDE, ta = tmem_base + (uint32_t)(j & 1) * 128 + q * EPT;
					const int nkk = (min(EPT, max(0, valid - q * EPT)) + 15) >> 4;
					#pragma unroll
					for (int kk = 0; kk < EPT / 16; ++kk)
						if (kk < nkk) umma_f16_ts(td, ta + kk * 8, bd + (uint64_t)(kk * 128), idesc_pv, (j | kk) ? 1u : 0u);
				}
				umma_commit(pv_full);
				umma_commit(&v_empty[s]);
				if (more) qk_mmas(j + 2);                // in order after the products: S[j & 1] / P may be overwritten
			}
			__syncwarp();
		}
	} else {
		// ===== softmax: warp = lane group (TMEM lanes) + 4 * key quarter; one (row, quarter) per thread =====
		const int kq = warp >> 2, lg = warp & 3;
		const int r = lg * 32 + lane;
		const uint32_t lane_off = (uint32_t)(lg * 32) << 16;
		const uint32_t ts0 = tmem_base + lane_off + kq * EPT;                    // my EPT scores in buffer 0 (+128 for buffer 1)
		const uint32_t to = tmem_base + O_BASE + kq * O_STRIDE + lane_off;       // my quarter's output accumulator of this row
		const float sl2 = p.scale_log2;
		float m = -INFINITY, l = 0.f;
		bool first = true;

		auto block = [&](int j, auto full_tag) {
			constexpr bool FULL = decltype(full_tag)::value;
			const int valid = FULL ? EPT : min(EPT, max(0, p.nk - j * AK - kq * EPT));      // valid keys of my part
			const uint32_t ts = ts0 + (uint32_t)(j & 1) * 128;
			if (p.park) mbar_wait_parked(&s_full[j & 1], (uint32_t)(j >> 1) & 1); else mbar_wait(&s_full[j & 1], (uint32_t)(j >> 1) & 1);
			tc_fence_after();
			uint32_t v[EPT];
			if (!FULL && valid <= 0) {                                           // nothing of this block belongs to my part
				#pragma unroll
				for (int i = 0; i < 16; ++i) v[i] = 0u;
				#pragma unroll
				for (int c = 0; c < EPT / 32; ++c) tmem_st16(ts + c * 16, v);
				tmem_st_wait(); tc_fence_before(); mbar_arrive(&p_full[j & 1]);
				return;
			}
			#pragma unroll
			for (int c = 0; c < EPT / 32; ++c) tmem_ld32(ts + c * 32, v + c * 32);
			tmem_ld_wait();
			float mx4[4] = { -INFINITY, -INFINITY, -INFINITY, -INFINITY };
			if (FULL) {
				#pragma unroll
				for (int i = 0; i < EPT; i += 8) {
					mx4[0] = max3f(mx4[0], __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
					mx4[1] = max3f(mx4[1], __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
					mx4[2] = max3f(mx4[2], __uint_as_float(v[i + 4]), __uint_as_float(v[i + 5]));
					mx4[3] = max3f(mx4[3], __uint_as_float(v[i + 6]), __uint_as_float(v[i + 7]));
				}
			} else {
				#pragma unroll
				for (int i = 0; i < EPT; ++i) if (i < valid) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(v[i]));
			}
			const float m_blk = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * sl2;
			// Lazy rescaling (as above): keep the old maximum while the block maximum exceeds it by < 2^8.
			if (first) { m = m_blk; first = false; }
			else {
				const bool need = m_blk > m + 8.0f;
				if (__any_sync(0xffffffffu, need)) {
					const float m_new = need ? m_blk : m;
					const float corr = ex2_approx(m - m_new);
					m = m_new;
					l *= corr;
					// PV(j-2) completed before QK(j) did (in-order tensor pipe), so the barrier's open phase is j-1 or later:
					// this waits exactly for PV(j-1), after which no product is in flight towards my accumulator
					mbar_wait(pv_full, (uint32_t)(j - 1) & 1);
					tc_fence_after();
					#pragma unroll
					for (int c0 = 0; c0 < 64; c0 += 16) {
						if (c0 < p.d16) {
							uint32_t o[16];
							tmem_ld16(to + c0, o);
							tmem_ld_wait();
							#pragma unroll
							for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * corr);
							tmem_st16(to + c0, o);
						}
					}
				}
			}
			const float mneg = -m;
			float rs4[4] = { 0.f, 0.f, 0.f, 0.f };
			uint32_t packed[EPT / 2];
			#pragma unroll
			for (int i = 0; i < EPT; i += 2) {
				float e0, e1;
				{ const float xs = fmaf(__uint_as_float(v[i]), sl2, mneg); e0 = (i & 7) < N_POLY ? ex2_poly(xs) : ex2_approx(xs); }
				{ const float xs = fmaf(__uint_as_float(v[i + 1]), sl2, mneg); e1 = ((i + 1) & 7) < N_POLY ? ex2_poly(xs) : ex2_approx(xs); }
				if (!FULL) { if (i >= valid) e0 = 0.f; if (i + 1 >= valid) e1 = 0.f; }
				rs4[(i >> 1) & 3] += e0 + e1;
				__half2 hh = __floats2half2_rn(e0, e1);
				packed[i >> 1] = *reinterpret_cast<uint32_t*>(&hh);
			}
			#pragma unroll
			for (int c = 0; c < EPT / 32; ++c) tmem_st16(ts + c * 16, packed + c * 16);
			l += (rs4[0] + rs4[1]) + (rs4[2] + rs4[3]);
			tmem_st_wait();
			tc_fence_before();
			mbar_arrive(&p_full[j & 1]);
		};
		const int nfull = p.nk / AK;      // a partial block can only be the last one
		for (int j = 0; j < nfull; ++j) block(j, std::true_type{});
		if (nfull < p.nblk) block(nfull, std::false_type{});

		// epilogue: merge the quarters. Every thread publishes (m, l) of its (row, quarter); quarter c takes the 16-column chunk c
		// of all four accumulators of its row.
		mbar_wait(pv_full, (uint32_t)(p.nblk - 1) & 1);
		tc_fence_after();
		exch[kq * 128 + r] = make_float2(first ? -INFINITY : m, l);
		named_bar_sync(1, 128 * NKQ);
		float mq[NKQ], lq[NKQ], M = -INFINITY;
		#pragma unroll
		for (int q = 0; q < NKQ; ++q) { const float2 e = exch[q * 128 + r]; mq[q] = e.x; lq[q] = e.y; M = fmaxf(M, e.x); }
		float wq[NKQ], L = 0.f;
		#pragma unroll
		for (int q = 0; q < NKQ; ++q) { wq[q] = (mq[q] == -INFINITY) ? 0.f : ex2_approx(mq[q] - M); L += lq[q] * wq[q]; }
		const float inv = L > 0.f ? 1.0f / L : 0.f;
		#pragma unroll
		for (int c0 = kq * 16; c0 < 64; c0 += NKQ * 16)       // 16-column chunk c of the output is merged by part c % NKQ
		if (c0 < p.d16) {
			const long long tok = (long long)q0 + r;
			float o[16];
			#pragma unroll
			for (int i = 0; i < 16; ++i) o[i] = 0.f;
			#pragma unroll
			for (int q = 0; q < NKQ; ++q) {
				uint32_t oq[16];
				tmem_ld16(tmem_base + O_BASE + q * O_STRIDE + lane_off + c0, oq);
				tmem_ld_wait();
				const float w = wq[q] * inv;
				#pragma unroll
				for (int i = 0; i < 16; ++i) o[i] = fmaf(__uint_as_float(oq[i]), w, o[i]);
			}
			if (tok < p.nq) {
				__half* op = (__half*)p.o + tok * p.so_t + (long long)h * p.so_h + (long long)b * p.so_b;
				const bool vec = ((((uintptr_t)op) & 15) == 0);
				#pragma unroll
				for (int h8 = 0; h8 < 16; h8 += 8) {
					const int cc = c0 + h8;
					if (cc < p.d) {
						if (vec && cc + 8 <= p.d) {
							uint4 o4; __half2* hp = reinterpret_cast<__half2*>(&o4);
							#pragma unroll
							for (int i = 0; i < 4; ++i) hp[i] = __floats2half2_rn(o[h8 + 2 * i], o[h8 + 2 * i + 1]);
							*reinterpret_cast<uint4*>(op + cc) = o4;
						} else {
							#pragma unroll
							for (int i = 0; i < 8; ++i) if (cc + i < p.d) op[cc + i] = __float2half_rn(o[h8 + i]);
						}
					}
				}
			}
		}
	}
	tc_fence_before();
	__syncthreads();
	if (warp == W_MMA) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}


// ------------------------------------------------------------------ heads up to 64 wide: two query tiles per CTA, scores three blocks deep
// What bounds the forms above (timelines in profiles/r2_attention_microbench.md, r3_attention.md): a softmax warp alternates between
// a phase without exponentials (wait for the scores, tcgen05.ld, row maximum, later pack + tcgen05.st + the round trip through the
// issuing warp and the tensor pipe: PV(j), QK(j+1)) and a phase that is nothing but exponentials. Warps that share a barrier run
// these phases together, so the MUFU pipe idles during the first and is oversubscribed during the second (45-60 % busy), and a
// single issuing thread that serves both tiles with predicated K-steps needs ~1500 clk per pair of blocks.
// Here ONE CTA per SM owns TWO tiles of 128 query rows (streams A, B: warps 0-3 / 4-7, one row per thread) and walks the keys in
// blocks of 64:
//   * the scores of a stream sit in a ring of THREE buffers (3 x 64 tensor-memory columns): QK(u+3) is issued right behind
//     PV(u), so a stream finds the scores of the next TWO blocks complete -- the tensor-pipe round trip is off the critical path;
//   * the softmax loop is SOFTWARE-PIPELINED inside the thread: while the exponentials of block u (MUFU-bound) are in flight, the
//     same instruction stream loads the scores of block u + 1 and takes their maximum, so a warp has no phase without exponentials
//     except the store / arrive tail;
//   * one issuing warp PER STREAM, straight-line issue (compile-time K-steps for d16 = 48 / 64), descriptors precomputed;
//   * packed f32x2 arithmetic, N_POLY pairs of every 8 pairs as polynomial on the FMA pipe;
//   * SUMMMA (d16 <= 48): the row sums come from the tensor core (P x ones, N = 16, into the 16 columns right of the accumulator).
// Tensor memory (512 columns): S_A[3] 0..191, S_B[3] 192..383, O_A 384..447, O_B 448..511.
// mbarrier parity waits only tell consecutive phases apart, and a stream may run ahead of the tensor pipe, so every block waits
// for PV(u-1) before it hands over P(u): the thread is never more than one completion away from the phase it asks for.
constexpr int AP_KST = 4, AP_VST = 4, AP_SB = 3;
#ifndef AP_DBG             // timing experiments (-DAP_DBG=n builds, results wrong): 1 no PV products, 2 no sum products, 4 no QK, 8 no exponentials
#define AP_DBG 0
#endif
#ifdef ATTN_AP_TRACE       // timeline probe build (tools/attn_trace.cu): the clock reads cost ~10 % and stay out of the product
#define AP_TR(role, j, ev) ATTN_TR(role, j, ev)
#else
#define AP_TR(role, j, ev) do { } while (0)
#endif
template <int N_POLY, bool SUMMMA, int NK16>
__global__ void __launch_bounds__(352, 1)
attn_ap_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
	const AttnParams p)
{
	extern __shared__ __align__(1024) uint8_t smem_raw[];
	uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
	constexpr int TB = CHUNK_BYTES;                  // d <= 64: one 128-byte chunk per row, 128 rows
	uint8_t* sQ = smem;                              // [2 tiles]
	uint8_t* sK = sQ + 2 * TB;                       // [AP_KST] tiles of 128 keys
	uint8_t* sV = sK + AP_KST * TB;                  // [AP_VST]
	uint8_t* sOnes = sV + AP_VST * TB;               // 16 key rows of 128 B, every f16 = 1.0 (any swizzle of it is itself)
	uint64_t* bars = (uint64_t*)(sOnes + 2048);
	uint64_t* q_full = bars;                         // [2]
	uint64_t* k_full = q_full + 2;                   // [AP_KST]
	uint64_t* k_empty = k_full + AP_KST;
	uint64_t* v_full = k_empty + AP_KST;             // [AP_VST]
	uint64_t* v_empty = v_full + AP_VST;
	uint64_t* s_full = v_empty + AP_VST;             // [2 streams][3 buffers]  QK done
	uint64_t* p_full = s_full + 2 * AP_SB;           // [2][3]  probabilities written (128 arrivals)
	uint64_t* pv_full = p_full + 2 * AP_SB;          // [2]     PV product of a block done
	uint32_t* tmem_slot = (uint32_t*)(pv_full + 2);

	const int warp = warp_id_uniform(), lane = threadIdx.x & 31;
	const int q0 = blockIdx.x * (2 * AQ), h = blockIdx.y, b = blockIdx.z;
	constexpr uint32_t S_STRIDE = 64 * AP_SB, O_BASE = 2 * S_STRIDE, O_STRIDE = 64;
	constexpr int W_TMA = 8, W_MMA = 9;              // warps 9 / 10 issue for stream A / B
	const int n64 = (p.nk + 63) >> 6;                // key blocks of 64 (>= 3: single-tile contexts run attn_kv1_kernel)
	const int nfull = p.nk >> 6;                     // a partial block can only be the last one

	if (threadIdx.x == 0) {
		tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
		for (int s = 0; s < AP_KST; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 2); }      // freed by both streams
		for (int s = 0; s < AP_VST; ++s) { mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 2); }
		for (int t = 0; t < 2; ++t) { mbar_init(&q_full[t], 1); mbar_init(&pv_full[t], 1); }
		for (int t = 0; t < 2 * AP_SB; ++t) { mbar_init(&s_full[t], 1); mbar_init(&p_full[t], 128); }
		fence_barrier_init();
	}
	if (SUMMMA) {
		if (threadIdx.x < 128) reinterpret_cast<uint4*>(sOnes)[threadIdx.x] = make_uint4(0x3C003C00u, 0x3C003C00u, 0x3C003C00u, 0x3C003C00u);
		fence_proxy_async();
	}
	if (warp == W_MMA) tmem_alloc(tmem_slot, 512);
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem_base = uniform_u32(*tmem_slot);

	if (warp == W_TMA) {
		if (lane == 0) {
			auto load_k = [&](int j) {
				const int s = j % AP_KST; const uint32_t ph = (uint32_t)(j / AP_KST) & 1;
				mbar_wait_parked(&k_empty[s], ph ^ 1);
				mbar_expect_tx(&k_full[s], TB);
				tma_load_4d(sK + (size_t)s * TB, &tmK, &k_full[s], 0, j * AK, h, b);
			};
			auto load_v = [&](int j) {
				const int s = j % AP_VST; const uint32_t ph = (uint32_t)(j / AP_VST) & 1;
				mbar_wait_parked(&v_empty[s], ph ^ 1);
				mbar_expect_tx(&v_full[s], TB);
				tma_load_4d(sV + (size_t)s * TB, &tmV, &v_full[s], 0, j * AK, h, b);
			};
			mbar_expect_tx(&q_full[0], TB);
			tma_load_4d(sQ, &tmQ, &q_full[0], 0, q0, h, b);
			load_k(0);
			mbar_expect_tx(&q_full[1], TB);
			tma_load_4d(sQ + TB, &tmQ, &q_full[1], 0, q0 + AQ, h, b);
			load_k(1);                                // nblk >= 2
			// K runs two tiles ahead of V: QK(u+3) is issued together with PV(u)
			for (int j = 0; j < p.nblk; ++j) { load_v(j); if (j + 2 < p.nblk) load_k(j + 2); }
		}
	} else if (warp >= W_MMA) {
		// ===== one issuing warp per stream; straight-line issue for full blocks =====
		const int t = warp - W_MMA;
		const int nk16 = NK16 ? NK16 : (p.d16 >> 4);
		const uint32_t idesc_qk = make_idesc_f16(AQ, 64, 0, 0);
		const uint32_t idesc_pv = make_idesc_f16(AQ, p.d16, 0, 1);     // A = P (TMEM), B = V MN-major ([key][d], d contiguous)
		const uint32_t idesc_sum = make_idesc_f16(AQ, 16, 0, 1);
		constexpr uint32_t tile16 = (uint32_t)TB >> 4;
		const uint64_t qdesc = make_smem_desc_sw128(smem_u32(sQ), 16, 1024) + (uint64_t)(t * tile16);
		const uint64_t kdesc0 = make_smem_desc_sw128(smem_u32(sK), 16, 1024);
		const uint64_t vdesc0 = make_smem_desc_sw128(smem_u32(sV), CHUNK_BYTES, 1024);
		const uint64_t odesc = make_smem_desc_sw128(smem_u32(sOnes), CHUNK_BYTES, 1024);
		const uint32_t t_s = tmem_base + (uint32_t)t * S_STRIDE, t_o = tmem_base + O_BASE + (uint32_t)t * O_STRIDE;
		int k_waited = -1, v_waited = -1;             // highest K / V tile seen resident
		// S_t[sb] = Q_t K[64 u .. 64 u + 64)^T : K tile u >> 1, rows 64 (u & 1).. (+8 KB = +512 in 16-byte units)
		auto qk = [&](int u, int sb) {
			const int j = u >> 1, ks = j % AP_KST;
			if (j > k_waited) { mbar_wait(&k_full[ks], (uint32_t)(j / AP_KST) & 1); k_waited = j; tc_fence_after(); }
			if (elect_one()) {
				const uint64_t bd = kdesc0 + (uint64_t)(ks * tile16) + (uint64_t)((u & 1) * 512);
				const uint32_t td = t_s + (uint32_t)(sb * 64);
				if (!(AP_DBG & 4) || u < 3) {
				#pragma unroll
				for (int kk = 0; kk < 4; ++kk)
					if (kk < nk16) umma_f16(td, qdesc + (uint64_t)(kk * 2), bd + (uint64_t)(kk * 2), idesc_qk, kk ? 1u : 0u);
				}
				umma_commit(&s_full[t * AP_SB + sb]);
				if ((u & 1) || u == n64 - 1) umma_commit(&k_empty[ks]);          // this stream is done with the K tile (2 arrivals free it)
			}
			__syncwarp();
		};
		// O_t += P_t(u) V[64 u ..): P = 32 packed columns over the start of S_t[sb]; 16 keys = 8 columns of P, 16 rows of V
		auto pv = [&](int u, int sb, auto full_tag) {
			constexpr bool FULL = decltype(full_tag)::value;
			const int j = u >> 1, vs = j % AP_VST;
			if (j > v_waited) { mbar_wait(&v_full[vs], (uint32_t)(j / AP_VST) & 1); v_waited = j; tc_fence_after(); }
			if (elect_one()) {
				const int nkk = FULL ? 4 : ((min(64, p.nk - u * 64) + 15) >> 4);
				const uint64_t bd = vdesc0 + (uint64_t)(vs * tile16) + (uint64_t)((u & 1) * 512);
				const uint32_t ta = t_s + (uint32_t)(sb * 64);
				const uint32_t acc = u ? 1u : 0u;
				if (!(AP_DBG & 1) || u < 1) {
				#pragma unroll
				for (int kk = 0; kk < 4; ++kk)
					if (FULL || kk < nkk) umma_f16_ts(t_o, ta + kk * 8, bd + (uint64_t)(kk * 128), idesc_pv, kk ? 1u : acc);
				}
				if (SUMMMA && (!(AP_DBG & 2) || u < 1)) {
					#pragma unroll
					for (int kk = 0; kk < 4; ++kk)
						if (FULL || kk < nkk) umma_f16_ts(t_o + p.d16, ta + kk * 8, odesc, idesc_sum, kk ? 1u : acc);
				}
				umma_commit(&pv_full[t]);
				if ((u & 1) || u == n64 - 1) umma_commit(&v_empty[vs]);
			}
			__syncwarp();
		};
		mbar_wait(&q_full[t], 0);
		tc_fence_after();
		qk(0, 0); qk(1, 1); qk(2, 2);                 // n64 >= 3
		int sb = 0; uint32_t ph = 0;                  // ring buffer and barrier phase of block u
		for (int u = 0; u < n64; ++u) {
			if (u + 3 < n64) {                        // K tile of QK(u+3): waited for before the probabilities arrive
				const int j = (u + 3) >> 1;
				if (j > k_waited) { mbar_wait(&k_full[j % AP_KST], (uint32_t)(j / AP_KST) & 1); k_waited = j; }
			}
			{ const int j = u >> 1; if (j > v_waited) { mbar_wait(&v_full[j % AP_VST], (uint32_t)(j / AP_VST) & 1); v_waited = j; } }
			mbar_wait(&p_full[t * AP_SB + sb], ph);
			tc_fence_after();
			if (lane == 0) AP_TR(2, u, t * 4 + 0);
			if (u < nfull) pv(u, sb, std::true_type{}); else pv(u, sb, std::false_type{});
			if (lane == 0) AP_TR(2, u, t * 4 + 1);
			if (u + 3 < n64) qk(u + 3, sb);           // in order behind PV(u): S_t[sb] / P_t(u) may be overwritten
			if (lane == 0) AP_TR(2, u, t * 4 + 2);
			if (++sb == AP_SB) { sb = 0; ph ^= 1; }
		}
	} else {
		// ===== softmax: stream t = warp / 4 (query tile), TMEM lane group = warp % 4; one query row per thread =====
		const int t = warp >> 2, quarter = warp & 3;
		const int r = quarter * 32 + lane;
		const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
		const uint32_t ts0 = tmem_base + lane_off + (uint32_t)t * S_STRIDE;
		const uint32_t to = tmem_base + O_BASE + (uint32_t)t * O_STRIDE + lane_off;
		const float sl2 = p.scale_log2;
		float m = -INFINITY, l = 0.f;
#ifdef ATTN_AP_TRACE
		const bool tr = p.trace != nullptr && quarter == 0 && lane == 0;
#else
		constexpr bool tr = false;
#endif

		auto row_max = [&](const uint32_t* va, const uint32_t* vb, int valid, auto full_tag) -> float {
			constexpr bool FULL = decltype(full_tag)::value;
			float mx4[4] = { -INFINITY, -INFINITY, -INFINITY, -INFINITY };
			if (FULL) {
				#pragma unroll
				for (int i = 0; i < 32; i += 8) {
					mx4[0] = max3f(mx4[0], __uint_as_float(va[i]), __uint_as_float(va[i + 1]));
					mx4[1] = max3f(mx4[1], __uint_as_float(va[i + 2]), __uint_as_float(va[i + 3]));
					mx4[2] = max3f(mx4[2], __uint_as_float(va[i + 4]), __uint_as_float(va[i + 5]));
					mx4[3] = max3f(mx4[3], __uint_as_float(va[i + 6]), __uint_as_float(va[i + 7]));
					mx4[0] = max3f(mx4[0], __uint_as_float(vb[i]), __uint_as_float(vb[i + 1]));
					mx4[1] = max3f(mx4[1], __uint_as_float(vb[i + 2]), __uint_as_float(vb[i + 3]));
					mx4[2] = max3f(mx4[2], __uint_as_float(vb[i + 4]), __uint_as_float(vb[i + 5]));
					mx4[3] = max3f(mx4[3], __uint_as_float(vb[i + 6]), __uint_as_float(vb[i + 7]));
				}
			} else {
				#pragma unroll
				for (int i = 0; i < 32; ++i) {
					if (i < valid) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(va[i]));
					if (32 + i < valid) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(vb[i]));
				}
			}
			return fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * sl2;
		};
		uint64_t acc0 = 0ull, acc1 = 0ull;
		// exponentials of 32 scores -> 16 packed columns of probabilities at tensor-memory address tp
		auto pass = [&](uint32_t* v, uint32_t tp, int c, int valid, uint64_t SL, uint64_t MN, auto full_tag) {
			constexpr bool FULL = decltype(full_tag)::value;
			uint32_t packed[16];
			#pragma unroll
			for (int i = 0; i < 32; i += 2) {
				const int pi = i >> 1;
				const uint64_t x2 = fma2(pk2u(v[i], v[i + 1]), SL, MN);
				float e0, e1;
				if (N_POLY > 0 && (((pi & 7) * N_POLY) & 7) < N_POLY) ex2_poly2(x2, e0, e1);
				else { float x0, x1; unpk2(x2, x0, x1); e0 = ex2_approx(x0); e1 = ex2_approx(x1); }
				if (!FULL) { if (c * 32 + i >= valid) e0 = 0.f; if (c * 32 + i + 1 >= valid) e1 = 0.f; }
				if (!SUMMMA) { if (pi & 1) acc1 = add2(acc1, pk2(e0, e1)); else acc0 = add2(acc0, pk2(e0, e1)); }
				__half2 hh = __floats2half2_rn(e0, e1);
				packed[pi] = *reinterpret_cast<uint32_t*>(&hh);
			}
			tmem_st16(tp, packed);
		};
		int sb = 0; uint32_t ph = 0;                  // ring buffer and barrier phase of block u
		// One iteration: block u sits in (ca, cb) with m already covering it; block u + 1 (if any) is loaded into (na, nb).
		auto step = [&](int u, uint32_t* ca, uint32_t* cb, uint32_t* na, uint32_t* nb, auto cur_full_tag, auto next_tag) {
			constexpr bool CUR_FULL = decltype(cur_full_tag)::value;
			constexpr int NEXT = decltype(next_tag)::value;                // 0 none, 1 full, 2 partial
			const int valid = CUR_FULL ? 64 : p.nk - u * 64;
			const uint32_t ts = ts0 + (uint32_t)sb * 64;
			int sbn = sb + 1; uint32_t phn = ph;
			if (sbn == AP_SB) { sbn = 0; phn ^= 1; }
			if (tr) AP_TR(t, u, 0);
			if (NEXT) {
				mbar_wait(&s_full[t * AP_SB + sbn], phn);
				tc_fence_after();
				const uint32_t tn = ts0 + (uint32_t)sbn * 64;
				tmem_ld32(tn, na); tmem_ld32(tn + 32, nb);
			}
			const uint64_t SL = pk2(sl2, sl2), MN = pk2(-m, -m);
			if (tr) AP_TR(t, u, 1);
			const bool do_exp = !(AP_DBG & 8);
			if (do_exp) pass(ca, ts, 0, valid, SL, MN, cur_full_tag);
			if (tr) AP_TR(t, u, 2);
			if (u > 0) mbar_wait(&pv_full[t], (uint32_t)(u - 1) & 1);     // PV(u-1): issued a block ago; keeps the phase count in step
			float m_blk = 0.f;
			if (NEXT) {
				tmem_ld_wait();
				#pragma unroll
				for (int i = 0; i < 32; ++i) { asm volatile("" : "+r"(na[i])); asm volatile("" : "+r"(nb[i])); }
				if (tr) AP_TR(t, u, 3);
				m_blk = NEXT == 1 ? row_max(na, nb, 64, std::true_type{}) : row_max(na, nb, p.nk - (u + 1) * 64, std::false_type{});
			}
			if (do_exp) pass(cb, ts + 16, 1, valid, SL, MN, cur_full_tag);
			if (tr) AP_TR(t, u, 4);
			tmem_st_wait();
			tc_fence_before();
			if (tr) AP_TR(t, u, 5);
			mbar_arrive(&p_full[t * AP_SB + sb]);
			if (NEXT) {
				// Lazy rescaling: keep the old maximum while the block maximum exceeds it by < 2^8 (f16 probabilities stay < 2^8 x 1).
				const bool need = m_blk > m + 8.0f;
				if (__any_sync(0xffffffffu, need)) {
					const float m_new = need ? m_blk : m;
					const float corr = ex2_approx(m - m_new);
					m = m_new;
					if (!SUMMMA) { float a0, a1; unpk2(add2(acc0, acc1), a0, a1); l = (l + a0 + a1) * corr; acc0 = 0ull; acc1 = 0ull; }
					mbar_wait(&pv_full[t], (uint32_t)u & 1);             // PV(u), issued behind the arrival above
					tc_fence_after();
					#pragma unroll
					for (int c0 = 0; c0 < 64; c0 += 16) {
						if (c0 < p.d16 + (SUMMMA ? 16 : 0)) {
							uint32_t o[16];
							tmem_ld16(to + c0, o);
							tmem_ld_wait();
							#pragma unroll
							for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * corr);
							tmem_st16(to + c0, o);
						}
					}
					tmem_st_wait();
				}
			}
			sb = sbn; ph = phn;
		};
		uint32_t a0r[32], b0r[32], a1r[32], b1r[32];
		{   // block 0
			mbar_wait_parked(&s_full[t * AP_SB], 0);
			tc_fence_after();
			tmem_ld32(ts0, a0r); tmem_ld32(ts0 + 32, b0r);
			tmem_ld_wait();
			m = row_max(a0r, b0r, 64, std::true_type{});          // n64 >= 3: block 0 is full
		}
		// blocks 0 .. n64 - 2 have a successor; the successor of block n64 - 2 is the last block (full or partial)
		const bool last_partial = nfull < n64;
		int u = 0;
		for (; u + 2 < n64 - 1; u += 2) {
			step(u, a0r, b0r, a1r, b1r, std::true_type{}, std::integral_constant<int, 1>{});
			step(u + 1, a1r, b1r, a0r, b0r, std::true_type{}, std::integral_constant<int, 1>{});
		}
		// here u is even and n64 - 1 - u is 1 or 2
		if (n64 - 1 - u == 2) {
			step(u, a0r, b0r, a1r, b1r, std::true_type{}, std::integral_constant<int, 1>{});
			if (last_partial) {
				step(u + 1, a1r, b1r, a0r, b0r, std::true_type{}, std::integral_constant<int, 2>{});
				step(u + 2, a0r, b0r, a1r, b1r, std::false_type{}, std::integral_constant<int, 0>{});
			} else {
				step(u + 1, a1r, b1r, a0r, b0r, std::true_type{}, std::integral_constant<int, 1>{});
				step(u + 2, a0r, b0r, a1r, b1r, std::true_type{}, std::integral_constant<int, 0>{});
			}
		} else {
			if (last_partial) {
				step(u, a0r, b0r, a1r, b1r, std::true_type{}, std::integral_constant<int, 2>{});
				step(u + 1, a1r, b1r, a0r, b0r, std::false_type{}, std::integral_constant<int, 0>{});
			} else {
				step(u, a0r, b0r, a1r, b1r, std::true_type{}, std::integral_constant<int, 1>{});
				step(u + 1, a1r, b1r, a0r, b0r, std::true_type{}, std::integral_constant<int, 0>{});
			}
		}
		if (!SUMMMA) { float s0, s1; unpk2(add2(acc0, acc1), s0, s1); l += s0 + s1; }

		// epilogue: O / l -> f16 -> global memory. PV(n64-2) was waited for by the last block: one completion to go.
		mbar_wait(&pv_full[t], (uint32_t)(n64 - 1) & 1);
		tc_fence_after();
		if (SUMMMA) {
			uint32_t t16[16];
			tmem_ld16(to + p.d16, t16);
			tmem_ld_wait();
			l = __uint_as_float(t16[0]);
		}
		const float inv = l > 0.f ? 1.0f / l : 0.f;
		const long long tok = (long long)q0 + t * AQ + r;
		__half* op = (__half*)p.o + tok * p.so_t + (long long)h * p.so_h + (long long)b * p.so_b;
		const bool vec = ((((uintptr_t)op) & 15) == 0);
		#pragma unroll
		for (int c0 = 0; c0 < 64; c0 += 16) {
			if (c0 < p.d16) {
				uint32_t oa[16];
				tmem_ld16(to + c0, oa);
				tmem_ld_wait();
				if (tok < p.nq) {
					#pragma unroll
					for (int h8 = 0; h8 < 16; h8 += 8) {
						const int cc = c0 + h8;
						if (cc < p.d) {
							if (vec && cc + 8 <= p.d) {
								uint4 o4; __half2* hp = reinterpret_cast<__half2*>(&o4);
								#pragma unroll
								for (int i = 0; i < 4; ++i) hp[i] = __floats2half2_rn(__uint_as_float(oa[h8 + 2 * i]) * inv, __uint_as_float(oa[h8 + 2 * i + 1]) * inv);
								*reinterpret_cast<uint4*>(op + cc) = o4;
							} else {
								#pragma unroll
								for (int i = 0; i < 8; ++i) if (cc + i < p.d) op[cc + i] = __float2half_rn(__uint_as_float(oa[h8 + i]) * inv);
							}
						}
					}
				}
			}
		}
	}
	tc_fence_before();
	__syncthreads();
	if (warp == W_MMA) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ------------------------------------------------------------------ single key block (cross-attention, nk <= 128)
// The text context of a cross-attention has 77 keys (unet.c:110-145): one key block. With one (pair of) query tile(s)
// per CTA such a launch is a chain of latencies -- load, QK^T, softmax, PV, store -- that nothing overlaps, ~6 us per
// CTA and 14 waves for the 64x64 level. Here a CTA keeps K and V of its (head, image) in shared memory and WALKS the
// query tiles: tile i uses slot i & 1 (its own Q buffer, score / output columns in tensor memory and softmax warp
// group), so the load, the two products and the softmax + store of consecutive tiles overlap.
//   warp 8  TMA producer: K, V once; Q tiles as their slot's buffer frees up (q_empty)
//   warp 9  MMA issuer: S_t = Q_t K^T (N = keys rounded up to 16), O_t = P_t V
//   warps 0..3 / 4..7  slot a / b: one query row per thread -- maximum, exponentials, P (f16, over the consumed scores),
//           then O / l -> f16 -> global memory, and the slot's accumulator is handed back (o_free).
template <int D16MAX>
__global__ void __launch_bounds__(320, D16MAX <= 64 ? 2 : 1)
attn_kv1_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
	const AttnParams p)
{
	extern __shared__ __align__(1024) uint8_t smem_raw[];
	uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
	const int tile_bytes = p.dchunks * CHUNK_BYTES;
	uint8_t* sQ = smem;                       // [2 slots]
	uint8_t* sK = sQ + 2 * tile_bytes;
	uint8_t* sV = sK + tile_bytes;
	uint64_t* bars = (uint64_t*)(sV + tile_bytes);
	uint64_t* kv_full = bars;                 // [1]
	uint64_t* q_full = kv_full + 1;           // [2]
	uint64_t* q_empty = q_full + 2;           // [2]
	uint64_t* s_full = q_empty + 2;           // [2]
	uint64_t* p_full = s_full + 2;            // [2]  128 arrivals
	uint64_t* pv_full = p_full + 2;           // [2]
	uint64_t* o_free = pv_full + 2;           // [2]  one arrival per warp of the slot
	uint32_t* tmem_slot = (uint32_t*)(o_free + 2);

	const int warp = warp_id_uniform(), lane = threadIdx.x & 31;
	const int h = blockIdx.y, b = blockIdx.z;
	const int n_tiles = (p.nq + AQ - 1) / AQ;
	const int n_my = ((int)blockIdx.x < n_tiles) ? (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;   // tiles blockIdx.x, + gridDim.x, ...
	// TMEM: S_t at t*128 (P_t over its first columns), O_t at 256 + t*128. COMPACT (the 77-token context with heads up to 48 wide:
	// SD1.x level 0): scores 80 + output 48 columns fit one 128-column slot, the CTA allocates 256 columns and TWO CTAs share an SM --
	// the kernel is a latency chain per tile (load, QK^T, softmax, PV, store), so twice the tiles in flight is what it needs
	const int nk16 = (p.nk + 15) & ~15;                   // keys the products cover (rows past nk are zero-filled by TMA)
	const bool compact = D16MAX <= 64 && nk16 <= 80 && p.d16 <= 48;
	const uint32_t O_BASE = compact ? 80u : 256u, O_STRIDE = 128, tmem_cols = compact ? 256u : (uint32_t)A_TMEM_COLS;

	if (threadIdx.x == 0) {
		tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
		mbar_init(kv_full, 1);
		for (int t = 0; t < 2; ++t) { mbar_init(&q_full[t], 1); mbar_init(&q_empty[t], 1); mbar_init(&s_full[t], 1); mbar_init(&p_full[t], 128);
			mbar_init(&pv_full[t], 1); mbar_init(&o_free[t], 4); }
		fence_barrier_init();
	}
	if (warp == 9) tmem_alloc(tmem_slot, tmem_cols);
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem_base = uniform_u32(*tmem_slot);

	if (warp == 8) {
		if (lane == 0 && n_my > 0) {
			mbar_expect_tx(kv_full, 2 * tile_bytes);
			for (int c = 0; c < p.dchunks; ++c) tma_load_4d(sK + c * CHUNK_BYTES, &tmK, kv_full, c * ACH, 0, h, b);
			for (int c = 0; c < p.dchunks; ++c) tma_load_4d(sV + c * CHUNK_BYTES, &tmV, kv_full, c * ACH, 0, h, b);
			for (int i = 0; i < n_my; ++i) {
				const int t = i & 1, n = i >> 1;
				if (n > 0) mbar_wait(&q_empty[t], (uint32_t)(n - 1) & 1);
				mbar_expect_tx(&q_full[t], tile_bytes);
				const int q0 = ((int)blockIdx.x + i * (int)gridDim.x) * AQ;
				for (int c = 0; c < p.dchunks; ++c) tma_load_4d(sQ + t * tile_bytes + c * CHUNK_BYTES, &tmQ, &q_full[t], c * ACH, q0, h, b);
			}
		}
	} else if (warp == 9) {
		const uint32_t idesc_qk = make_idesc_f16(AQ, nk16, 0, 0);
		const uint32_t idesc_pv = make_idesc_f16(AQ, p.d16, 0, 1);
		const uint64_t qdesc0 = make_smem_desc_sw128(smem_u32(sQ), 16, 1024);
		const uint64_t kdesc0 = make_smem_desc_sw128(smem_u32(sK), 16, 1024);
		const uint64_t vdesc0 = make_smem_desc_sw128(smem_u32(sV), CHUNK_BYTES, 1024);
		const uint32_t tile16 = (uint32_t)tile_bytes >> 4;
		const int nd16 = p.d16 >> 4, nkk = nk16 >> 4;
		auto issue_qk = [&](int i) {
			const int t = i & 1, n = i >> 1;
			mbar_wait(&q_full[t], (uint32_t)n & 1);
			tc_fence_after();
			if (elect_one()) {
				const uint64_t ad = qdesc0 + (uint64_t)(t * tile16);
				const uint32_t td = tmem_base + t * 128;
				#pragma unroll
				for (int kk = 0; kk < D16MAX / 16; ++kk) {
					const uint32_t off = (uint32_t)((kk >> 2) * (CHUNK_BYTES >> 4) + (kk & 3) * 2);
					if (kk < nd16) umma_f16(td, ad + off, kdesc0 + off, idesc_qk, kk ? 1u : 0u);
				}
				umma_commit(&s_full[t]);
				umma_commit(&q_empty[t]);
			}
			__syncwarp();
		};
		auto issue_pv = [&](int i) {
			const int t = i & 1, n = i >> 1;
			mbar_wait(&p_full[t], (uint32_t)n & 1);
			if (n > 0) mbar_wait(&o_free[t], (uint32_t)(n - 1) & 1);      // the slot's previous output has been read
			tc_fence_after();
			if (elect_one()) {
				const uint32_t td = tmem_base + O_BASE + t * O_STRIDE, ta = tmem_base + t * 128;
				#pragma unroll
				for (int kk = 0; kk < AK / 16; ++kk)
					if (kk < nkk) umma_f16_ts(td, ta + kk * 8, vdesc0 + (uint64_t)(kk * 128), idesc_pv, kk ? 1u : 0u);
				umma_commit(&pv_full[t]);
			}
			__syncwarp();
		};
		if (n_my > 0) {
			mbar_wait(kv_full, 0);
			issue_qk(0);
			if (n_my > 1) issue_qk(1);
			for (int i = 0; i < n_my; ++i) {
				issue_pv(i);                                   // in order after it, S_t / P_t may be overwritten
				if (i + 2 < n_my) issue_qk(i + 2);
			}
		}
	} else {
		const int t = warp >> 2, quarter = warp & 3;
		const int r = quarter * 32 + lane;
		const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
		const uint32_t ts = tmem_base + t * 128 + lane_off;
		const uint32_t to = tmem_base + O_BASE + t * O_STRIDE + lane_off;
		const float sl2 = p.scale_log2;
		const int nch = (p.nk + 31) >> 5;                     // 32-column chunks holding valid keys
		for (int n = 0; 2 * n + t < n_my; ++n) {
			const int i = 2 * n + t;
			mbar_wait(&s_full[t], (uint32_t)n & 1);
			tc_fence_after();
			uint32_t v[32];
			float mx = -INFINITY;
			for (int c = 0; c < nch; ++c) {
				tmem_ld32(ts + c * 32, v);
				tmem_ld_wait();
				#pragma unroll
				for (int k = 0; k < 32; ++k) if (c * 32 + k < p.nk) mx = fmaxf(mx, __uint_as_float(v[k]));
			}
			const float mneg = -mx * sl2;
			float l = 0.f;
			for (int c = 0; c < nch; ++c) {
				tmem_ld32(ts + c * 32, v);
				tmem_ld_wait();
				uint32_t packed[16];
				#pragma unroll
				for (int k = 0; k < 32; k += 2) {
					float p0 = ex2_approx(fmaf(__uint_as_float(v[k]), sl2, mneg)), p1 = ex2_approx(fmaf(__uint_as_float(v[k + 1]), sl2, mneg));
					if (c * 32 + k >= p.nk) p0 = 0.f;
					if (c * 32 + k + 1 >= p.nk) p1 = 0.f;
					l += p0 + p1;
					__half2 hh = __floats2half2_rn(p0, p1);
					packed[k >> 1] = *reinterpret_cast<uint32_t*>(&hh);
				}
				tmem_st16(ts + c * 16, packed);               // columns [16c, 16c+16) lie inside scores that were already read
			}
			tmem_st_wait();
			tc_fence_before();
			mbar_arrive(&p_full[t]);

			mbar_wait(&pv_full[t], (uint32_t)n & 1);
			tc_fence_after();
			const long long tok = ((long long)blockIdx.x + (long long)i * gridDim.x) * AQ + r;
			const float inv = l > 0.f ? 1.0f / l : 0.f;
			__half* op = (__half*)p.o + tok * p.so_t + (long long)h * p.so_h + (long long)b * p.so_b;
			const bool vec = ((((uintptr_t)op) & 15) == 0);
			#pragma unroll
			for (int c0 = 0; c0 < D16MAX; c0 += 16) {
				if (c0 < p.d16) {
					uint32_t o[16];
					tmem_ld16(to + c0, o);
					tmem_ld_wait();
					if (tok < p.nq) {
						#pragma unroll
						for (int h8 = 0; h8 < 16; h8 += 8) {
							const int cc = c0 + h8;
							if (cc < p.d) {
								if (vec && cc + 8 <= p.d) {
									uint4 o4; __half2* hp = reinterpret_cast<__half2*>(&o4);
									#pragma unroll
									for (int q = 0; q < 4; ++q) hp[q] = __floats2half2_rn(__uint_as_float(o[h8 + 2 * q]) * inv, __uint_as_float(o[h8 + 2 * q + 1]) * inv);
									*reinterpret_cast<uint4*>(op + cc) = o4;
								} else {
									#pragma unroll
									for (int q = 0; q < 8; ++q) if (cc + q < p.d) op[cc + q] = __float2half_rn(__uint_as_float(o[h8 + q]) * inv);
								}
							}
						}
					}
				}
			}
			tc_fence_before();
			__syncwarp();
			if (lane == 0) mbar_arrive(&o_free[t]);
		}
	}
	tc_fence_before();
	__syncthreads();
	if (warp == 9) { tc_fence_after(); tmem_dealloc(tmem_base, tmem_cols); }
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
	p.trace = nullptr;
	{ const char* e = getenv("GGML_B200_ATTN_PP"); p.pingpong = e ? atoi(e) : 1; }
	{ const char* e = getenv("GGML_B200_ATTN_SEP"); p.sep_p = e ? atoi(e) : 0; }
	{ const char* e = getenv("GGML_B200_ATTN_PARK"); p.park = e ? atoi(e) : 1; }       // long waits parked with a suspend-time hint (0: polled)
	// dual form: one exponential in eight on the FMA pipe (measured -4 % at d = 40, -2.6 % at d = 64; two in eight is worse)
	p.d = (int)q.ne[0]; p.d16 = (p.d + 15) / 16 * 16;
	{ const char* e = getenv("GGML_B200_ATTN_POLY"); p.npoly = e ? atoi(e) : (p.d16 > 48 && p.d16 <= 64 ? 0 : 1); }
	p.d16 = (p.d + 15) / 16 * 16; p.dchunks = (p.d + ACH - 1) / ACH;
	p.nq = (int)q.ne[1]; p.nk = (int)k.ne[1]; p.H = (int)q.ne[2]; p.B = (int)q.ne[3];
	p.nblk = (p.nk + AK - 1) / AK;
	p.scale_log2 = scale * 1.4426950408889634f;
	p.o = o.ptr; p.so_t = o.st[1]; p.so_h = o.st[2]; p.so_b = o.st[3];
	const size_t tile = (size_t)p.dchunks * CHUNK_BYTES;
	p.stages = std::max(1, std::min(p.nblk, A_MAX_STAGES));
	// heads up to 64 wide: one tile per CTA, two CTAs per SM (measured 7-12 % faster than two ping-pong tiles in one CTA)
	{ const char* e = getenv("GGML_B200_ATTN_DUAL"); p.dual = (e ? atoi(e) : 1) && p.d16 <= 64 && p.nblk > 1; }
	// key-split form of the dual layout (eight softmax warps per CTA): default for heads up to 64 wide
	// GGML_B200_ATTN_SPLIT: 0 = one row per thread (round-1 dual form), 2 = key halves, two CTAs per SM (default: 837 / 272 / 340 us on
	// the three shapes of profiles/r2_attention_microbench.md against 841 / 300 / 392 us), 4 = key quarters + double-buffered
	// scores, one CTA per SM (measured slower: 911 / 301 / 350 us -- kept selectable)
	// 5 = two query tiles per CTA, scores three blocks deep, software-pipelined softmax (attn_ap_kernel): default for heads up to 48
	// wide (SD1.x level 0: 650 us against 820-840 us); 64-wide heads stay on the key halves (253 / 311 us against 247 / 321 us)
	{ const char* e = getenv("GGML_B200_ATTN_SPLIT"); p.split = p.dual ? (e ? atoi(e) : (p.d16 <= 48 ? 5 : 2)) : 0; if (p.split && p.split != 4 && p.split != 3 && p.split != 5) p.split = 2; }
	const int nt = (p.d16 > 128 || (p.dual && p.split != 5)) ? 1 : 2;
	if (p.dual && p.split != 4 && p.split != 3 && p.split != 5) p.stages = std::min(p.stages, 2);             // two CTAs per SM: <= 113 KB each
	auto total = [&]() { return tile * (nt + 2 * p.stages) + 1024 + 512; };
	while (total() > 220 * 1024 && p.stages > 1) p.stages--;
	a->smem = total() + (p.split >= 3 ? 4096 : p.split ? 2048 : 0);       // + the (maximum, sum) exchange of the key-split forms
	// packed softmax arithmetic of the key-halves form (GGML_B200_ATTN_PK: 0 scalar, 1 packed, 2 packed + row sums on the tensor core)
	{
		const char* e = getenv("GGML_B200_ATTN_PK");
		p.pk = p.split == 2 ? (e ? atoi(e) : 1) : 0;     // packed: 276 -> 253 us, 341 -> 305 us on the 64-wide shapes
		if (p.pk == 2 && p.d16 > 48) p.pk = 1;
		if (p.pk == 2) a->smem += 2048;
		if (p.pk) { const char* ep = getenv("GGML_B200_ATTN_POLY"); p.npoly = std::max(0, std::min(2, ep ? atoi(ep) : 1)); }
	}
	if (p.split == 5) {          // two query tiles per CTA, scores three blocks deep (attn_ap_kernel)
		const char* e = getenv("GGML_B200_ATTN_PK"); p.pk = e ? atoi(e) : 2;
		if (p.pk != 1) p.pk = 2;
		if (p.d16 > 48) p.pk = 1;            // no tensor-memory columns left for the sum product next to a 64-wide accumulator
		const char* ep = getenv("GGML_B200_ATTN_POLY"); p.npoly = std::max(0, std::min(2, ep ? atoi(ep) : 1));
		a->smem = (size_t)(2 + AP_KST + AP_VST) * CHUNK_BYTES + 2048 + 1024 + 1024;
	}
	a->grid = dim3((unsigned)((p.nq + nt * AQ - 1) / (nt * AQ)), (unsigned)p.H, (unsigned)p.B);
	{
		// one key block (cross-attention): CTAs walk the query tiles of their (head, image); as many CTAs per (head, image)
		// as fit the chip in one wave
		const char* e = getenv("GGML_B200_ATTN_KV1");
		if (p.nblk == 1 && p.d16 <= 128 && !(e && atoi(e) == 0)) {
			const int n_tiles = (p.nq + AQ - 1) / AQ;
			const long long hb = (long long)p.H * p.B;
			const bool compact = p.d16 <= 48 && ((p.nk + 15) & ~15) <= 80;         // two CTAs per SM (256 tensor-memory columns each)
			int x = (int)std::max<long long>(1, (compact ? 296 : 148) / std::max<long long>(1, hb));
			x = std::min(x, n_tiles);
			a->kv1 = true;
			a->grid = dim3((unsigned)x, (unsigned)p.H, (unsigned)p.B);
			a->smem = tile * 4 + 1024 + 512;
		}
	}
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
		CUDA_CHECK(cudaFuncSetAttribute(attn_tc_kernel<64, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024));
		CUDA_CHECK(cudaFuncSetAttribute((attn_tc_kernel<64, 1, 1>), cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024));
		CUDA_CHECK(cudaFuncSetAttribute((attn_tc_kernel<64, 1, 2>), cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024));
		CUDA_CHECK(cudaFuncSetAttribute((attn_split4_kernel<0, 4>), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
		CUDA_CHECK(cudaFuncSetAttribute((attn_split4_kernel<1, 4>), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
		CUDA_CHECK(cudaFuncSetAttribute((attn_split4_kernel<2, 4>), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
		CUDA_CHECK(cudaFuncSetAttribute((attn_split4_kernel<0, 2>), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
		CUDA_CHECK(cudaFuncSetAttribute((attn_split4_kernel<1, 2>), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
		CUDA_CHECK(cudaFuncSetAttribute((attn_split_kernel<0, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024));
		CUDA_CHECK(cudaFuncSetAttribute((attn_split_kernel<1, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024));
		CUDA_CHECK(cudaFuncSetAttribute((attn_split_kernel<2, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024));
		CUDA_CHECK(cudaFuncSetAttribute((attn_split_kernel<0, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024));
		CUDA_CHECK(cudaFuncSetAttribute((attn_split_kernel<1, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024));
		CUDA_CHECK(cudaFuncSetAttribute((attn_split_kernel<2, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024));
		#define PK_ATTR(NP) \
			CUDA_CHECK(cudaFuncSetAttribute((attn_split_kernel<NP, false, 1>), cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024)); \
			CUDA_CHECK(cudaFuncSetAttribute((attn_split_kernel<NP, true, 1>), cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024)); \
			CUDA_CHECK(cudaFuncSetAttribute((attn_split_kernel<NP, false, 2>), cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024));
		PK_ATTR(0) PK_ATTR(1) PK_ATTR(2)
		#undef PK_ATTR
		#define AP_ATTR(NP) \
			CUDA_CHECK(cudaFuncSetAttribute((attn_ap_kernel<NP, true, 0>), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
			CUDA_CHECK(cudaFuncSetAttribute((attn_ap_kernel<NP, true, 3>), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
			CUDA_CHECK(cudaFuncSetAttribute((attn_ap_kernel<NP, false, 0>), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
			CUDA_CHECK(cudaFuncSetAttribute((attn_ap_kernel<NP, false, 3>), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
			CUDA_CHECK(cudaFuncSetAttribute((attn_ap_kernel<NP, false, 4>), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
		AP_ATTR(0) AP_ATTR(1) AP_ATTR(2)
		#undef AP_ATTR
		CUDA_CHECK(cudaFuncSetAttribute(attn_kv1_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
		CUDA_CHECK(cudaFuncSetAttribute(attn_kv1_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
		attr_set = true;
	}
	if (a->kv1) {
		if (a->p.d16 <= 64) attn_kv1_kernel<64><<<a->grid, 320, a->smem, s>>>(a->tmQ, a->tmK, a->tmV, a->p);
		else attn_kv1_kernel<128><<<a->grid, 320, a->smem, s>>>(a->tmQ, a->tmK, a->tmV, a->p);
		g_stats.kernel_launches++;
		return;
	}
	if (a->p.split == 5) {
		#define AP_K(NP, SM, K16) attn_ap_kernel<NP, SM, K16><<<a->grid, 352, a->smem, s>>>(a->tmQ, a->tmK, a->tmV, a->p)
		#define AP_CASE(NP) do { \
			if (a->p.pk == 2) { if (a->p.d16 == 48) AP_K(NP, true, 3); else AP_K(NP, true, 0); } \
			else { if (a->p.d16 == 48) AP_K(NP, false, 3); else if (a->p.d16 == 64) AP_K(NP, false, 4); else AP_K(NP, false, 0); } } while (0)
		switch (a->p.npoly) { case 0: AP_CASE(0); break; case 1: AP_CASE(1); break; default: AP_CASE(2); break; }
		#undef AP_CASE
		#undef AP_K
		g_stats.kernel_launches++;
		return;
	}
	if (a->p.split == 4) {
		if (a->p.npoly == 2) attn_split4_kernel<2, 4><<<a->grid, 576, a->smem, s>>>(a->tmQ, a->tmK, a->tmV, a->p);
		else if (a->p.npoly == 1) attn_split4_kernel<1, 4><<<a->grid, 576, a->smem, s>>>(a->tmQ, a->tmK, a->tmV, a->p);
		else attn_split4_kernel<0, 4><<<a->grid, 576, a->smem, s>>>(a->tmQ, a->tmK, a->tmV, a->p);
		g_stats.kernel_launches++;
		return;
	}
	if (a->p.split == 3) {          // key halves + double-buffered scores, one CTA per SM
		if (a->p.npoly >= 1) attn_split4_kernel<1, 2><<<a->grid, 320, a->smem, s>>>(a->tmQ, a->tmK, a->tmV, a->p);
		else attn_split4_kernel<0, 2><<<a->grid, 320, a->smem, s>>>(a->tmQ, a->tmK, a->tmV, a->p);
		g_stats.kernel_launches++;
		return;
	}
	if (a->p.split) {
		// measured (profiles/r2_attention_microbench.md): 64-wide heads gain 4-6 % from the anti-phase halves with all exponentials on
		// the MUFU pipe; 40-wide heads (one MMA fewer per product) are 1 % faster in phase with one exponential in eight on the FMA pipe
		static const int stag_env = getenv("GGML_B200_ATTN_STAGGER") ? atoi(getenv("GGML_B200_ATTN_STAGGER")) : -1;
		const bool stag = stag_env >= 0 ? stag_env != 0 : a->p.d16 > 48;
		if (a->p.pk) {
			#define PK_CASE(NP, ST, PKV) attn_split_kernel<NP, ST, PKV><<<a->grid, 320, a->smem, s>>>(a->tmQ, a->tmK, a->tmV, a->p)
			#define PK_NP(ST, PKV) switch (a->p.npoly) { case 0: PK_CASE(0, ST, PKV); break; case 1: PK_CASE(1, ST, PKV); break; \
				default: PK_CASE(2, ST, PKV); break; }
			if (a->p.pk == 2) { PK_NP(false, 2) }
			else if (stag) { PK_NP(true, 1) }
			else { PK_NP(false, 1) }
			#undef PK_NP
			#undef PK_CASE
		} else if (stag) {
			if (a->p.npoly == 2) attn_split_kernel<2, true><<<a->grid, 320, a->smem, s>>>(a->tmQ, a->tmK, a->tmV, a->p);
			else if (a->p.npoly == 1) attn_split_kernel<1, true><<<a->grid, 320, a->smem, s>>>(a->tmQ, a->tmK, a->tmV, a->p);
			else attn_split_kernel<0, true><<<a->grid, 320, a->smem, s>>>(a->tmQ, a->tmK, a->tmV, a->p);
		} else {
			if (a->p.npoly == 2) attn_split_kernel<2, false><<<a->grid, 320, a->smem, s>>>(a->tmQ, a->tmK, a->tmV, a->p);
			else if (a->p.npoly == 1) attn_split_kernel<1, false><<<a->grid, 320, a->smem, s>>>(a->tmQ, a->tmK, a->tmV, a->p);
			else attn_split_kernel<0, false><<<a->grid, 320, a->smem, s>>>(a->tmQ, a->tmK, a->tmV, a->p);
		}
		g_stats.kernel_launches++;
		return;
	}
	if (a->p.dual && a->p.npoly == 2) attn_tc_kernel<64, 1, 2><<<a->grid, 192, a->smem, s>>>(a->tmQ, a->tmK, a->tmV, a->p);
	else if (a->p.dual && a->p.npoly == 1) attn_tc_kernel<64, 1, 1><<<a->grid, 192, a->smem, s>>>(a->tmQ, a->tmK, a->tmV, a->p);
	else if (a->p.dual) attn_tc_kernel<64, 1><<<a->grid, 192, a->smem, s>>>(a->tmQ, a->tmK, a->tmV, a->p);
	else if (a->p.d16 <= 64) attn_tc_kernel<64, 2><<<a->grid, 320, a->smem, s>>>(a->tmQ, a->tmK, a->tmV, a->p);
	else if (a->p.d16 <= 128) attn_tc_kernel<128, 2><<<a->grid, 320, a->smem, s>>>(a->tmQ, a->tmK, a->tmV, a->p);
	else attn_tc_kernel<160, 1><<<a->grid, 192, a->smem, s>>>(a->tmQ, a->tmK, a->tmV, a->p);
	g_stats.kernel_launches++;
}

void attn_tc_free(AttnTC* a) { delete a; }
void attn_tc_set_trace(AttnTC* a, long long* dev_buf) { a->p.trace = dev_buf; }

}  // namespace b200
