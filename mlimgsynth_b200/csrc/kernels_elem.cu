// kernels_elem.cu -- generic strided kernels and the fused memory-bound kernels
// (GroupNorm+SiLU, LayerNorm, GEGLU gate, im2col, weight prep) of the B200 engine.
// Rooflines: everything in this file is HBM/L2-bandwidth bound; the hot ones are vectorised
// to 16-byte accesses with channels innermost so that a warp touches contiguous 512 B.
#include "kernels.h"
#include "tc_common.cuh"
#include <algorithm>

namespace b200 {

// ------------------------------------------------------------------ helpers
struct V4 {               // device copy of a View with an iteration order
	void* ptr; int dt;
	long long ne[4], st[4];
};
struct Iter4 { long long n[4]; int ord[4]; long long total; };

static V4 v4(const View& v) { V4 r; r.ptr = v.ptr; r.dt = v.dt; for (int i = 0; i < 4; ++i) { r.ne[i] = v.ne[i]; r.st[i] = v.st[i]; } return r; }

// Iterate dst in order of increasing dst stride so consecutive threads write consecutive memory.
static Iter4 iter_for(const View& dst)
{
	Iter4 it;
	int ord[4] = {0, 1, 2, 3};
	std::stable_sort(ord, ord + 4, [&](int a, int b) {
		long long sa = dst.ne[a] == 1 ? (1LL << 60) : dst.st[a], sb = dst.ne[b] == 1 ? (1LL << 60) : dst.st[b];
		return sa < sb; });
	it.total = 1;
	for (int i = 0; i < 4; ++i) { it.ord[i] = ord[i]; it.n[i] = dst.ne[ord[i]]; it.total *= it.n[i]; }
	return it;
}

__device__ __forceinline__ void decompose(const Iter4& it, long long lin, long long idx[4])
{
	#pragma unroll
	for (int i = 0; i < 4; ++i) { long long n = it.n[i]; long long q = lin / n; idx[it.ord[i]] = lin - q * n; lin = q; }
}
__device__ __forceinline__ float ldv(const V4& v, long long off)
{
	if (v.dt == DT_F16) return __half2float(((const __half*)v.ptr)[off]);
	if (v.dt == DT_I32) return (float)((const int*)v.ptr)[off];
	return ((const float*)v.ptr)[off];
}
__device__ __forceinline__ void stv(const V4& v, long long off, float x)
{
	if (v.dt == DT_F16) ((__half*)v.ptr)[off] = __float2half_rn(x);
	else if (v.dt == DT_I32) ((int*)v.ptr)[off] = (int)x;
	else ((float*)v.ptr)[off] = x;
}
__device__ __forceinline__ long long offs(const V4& v, const long long i[4])
{ return i[0] * v.st[0] + i[1] * v.st[1] + i[2] * v.st[2] + i[3] * v.st[3]; }

static inline int grid_for(long long total, int block, long long per_thread = 1)
{
	long long g = (total + (long long)block * per_thread - 1) / ((long long)block * per_thread);
	return (int)std::min<long long>(std::max<long long>(g, 1), 148LL * 32);
}

__device__ __forceinline__ float act_apply(int op, float x, float p)
{
	switch (op) {
	case U_TANH: return tanhf(x);
	case U_RELU: return fmaxf(x, 0.0f);
	case U_SILU: return x / (1.0f + __expf(-x));
	case U_GELU: {  // tanh approximation, as the reference's ggml_gelu
		float u = 0.79788456080286535588f * x * (1.0f + 0.044715f * x * x);
		return 0.5f * x * (1.0f + tanhf(u));
	}
	case U_GELU_QUICK: return x / (1.0f + __expf(-1.702f * x));
	case U_SCALE: return x * p;
	default: return x;
	}
}

// ------------------------------------------------------------------ copy / convert
__global__ void copy_kernel(V4 dst, V4 src, Iter4 it)
{
	for (long long lin = blockIdx.x * (long long)blockDim.x + threadIdx.x; lin < it.total;
	     lin += (long long)gridDim.x * blockDim.x) {
		long long i[4]; decompose(it, lin, i);
		if (dst.dt == DT_I32 && src.dt == DT_I32) ((int*)dst.ptr)[offs(dst, i)] = ((const int*)src.ptr)[offs(src, i)];
		else stv(dst, offs(dst, i), ldv(src, offs(src, i)));
	}
}
// Fast path (concat halves, pad interiors, nearest upscale, layout-preserving copies): both views have the same
// unit-stride dim whose extent, the other strides and both base pointers are multiples of 16 bytes. One 16-byte
// vector per thread step; src index = dst index / f (f = 1 for plain copies, the scale factors for upscale).
struct VecCopy {
	const uint4* src; uint4* dst;
	int n[4];                  // extents in iteration order, n[0] = vectors along the contiguous dim
	long long ds[4], ss[4];    // strides in vectors
	int f[4];                  // src index divisor per iteration dim
	long long total;
};
// Exact unsigned 32-bit division by an invariant divisor (Granlund-Montgomery): q = (t + ((n - t) >> s1)) >> s2 with
// t = umulhi(n, m). The index decomposition of the copy kernels would otherwise spend several hundred instructions per
// 16-byte vector on 64-bit divisions and be instruction-bound at ~2.8 TB/s.
struct FastDiv32 { unsigned m, s1, s2; };
static FastDiv32 fastdiv32(unsigned d)
{
	unsigned l = 0; while ((1ull << l) < d) ++l;
	FastDiv32 f; f.m = (unsigned)((((1ull << 32) * ((1ull << l) - d)) / d) + 1); f.s1 = l < 1 ? l : 1; f.s2 = l < 1 ? 0 : l - 1;
	return f;
}
__device__ __forceinline__ unsigned fd_div(unsigned n, const FastDiv32& f) { const unsigned t = __umulhi(n, f.m); return (t + ((n - t) >> f.s1)) >> f.s2; }

struct VecCopy32 {
	const uint4* src; uint4* dst;
	unsigned n[3]; FastDiv32 dn[3];      // extents of the three fastest iteration dims and their reciprocals
	long long ds[4], ss[4];
	unsigned fs[4];                       // log2 of the src index divisor (nearest upscale by 1 or 2)
	unsigned total;
};
__global__ void __launch_bounds__(256)
vec_copy32_kernel(VecCopy32 c)
{
	const unsigned stride = gridDim.x * blockDim.x;
	for (unsigned lin = blockIdx.x * blockDim.x + threadIdx.x; lin < c.total; lin += stride) {
		const unsigned r1 = fd_div(lin, c.dn[0]), i0 = lin - r1 * c.n[0];
		const unsigned r2 = fd_div(r1, c.dn[1]), i1 = r1 - r2 * c.n[1];
		const unsigned i3 = fd_div(r2, c.dn[2]), i2 = r2 - i3 * c.n[2];
		c.dst[i0 * c.ds[0] + i1 * c.ds[1] + i2 * c.ds[2] + i3 * c.ds[3]] =
			__ldg(c.src + i0 * c.ss[0] + (i1 >> c.fs[1]) * c.ss[1] + (i2 >> c.fs[2]) * c.ss[2] + (i3 >> c.fs[3]) * c.ss[3]);
	}
}

__global__ void vec_copy_kernel(VecCopy c)
{
	for (long long lin = blockIdx.x * (long long)blockDim.x + threadIdx.x; lin < c.total; lin += (long long)gridDim.x * blockDim.x) {
		long long r = lin;
		const int i0 = (int)(r % c.n[0]); r /= c.n[0];
		const int i1 = (int)(r % c.n[1]); r /= c.n[1];
		const int i2 = (int)(r % c.n[2]); const int i3 = (int)(r / c.n[2]);
		c.dst[i0 * c.ds[0] + i1 * c.ds[1] + i2 * c.ds[2] + i3 * c.ds[3]] =
			__ldg(c.src + i0 * c.ss[0] + (i1 / c.f[1]) * c.ss[1] + (i2 / c.f[2]) * c.ss[2] + (i3 / c.f[3]) * c.ss[3]);
	}
}
static bool try_vec_copy(cudaStream_t s, const View& dst, const View& src)
{
	if (dst.dt != src.dt || dst.dt == DT_I32) return false;
	const int es = (int)dt_size(dst.dt), vl = 16 / es;
	int d0 = -1;
	for (int i = 0; i < 4; ++i) if (dst.ne[i] > 1 && dst.st[i] == 1 && src.st[i] == 1 && dst.ne[i] == src.ne[i]) { d0 = i; break; }
	if (d0 < 0 || dst.ne[d0] % vl || ((uintptr_t)dst.ptr & 15) || ((uintptr_t)src.ptr & 15)) return false;
	int f[4];
	for (int i = 0; i < 4; ++i) {
		if (src.ne[i] <= 0 || dst.ne[i] % src.ne[i]) return false;
		f[i] = (int)(dst.ne[i] / src.ne[i]);
		if (i != d0 && ((dst.ne[i] > 1 && dst.st[i] % vl) || (src.ne[i] > 1 && src.st[i] % vl))) return false;
		if (dst.ne[i] > 0x7fffffff) return false;
	}
	// iteration order: contiguous dim first, then the others by increasing dst stride
	int ord[4], k = 0; ord[k++] = d0;
	for (int i = 0; i < 4; ++i) if (i != d0) ord[k++] = i;
	std::stable_sort(ord + 1, ord + 4, [&](int a, int b) {
		long long sa = dst.ne[a] == 1 ? (1LL << 60) : dst.st[a], sb = dst.ne[b] == 1 ? (1LL << 60) : dst.st[b]; return sa < sb; });
	VecCopy c; c.src = (const uint4*)src.ptr; c.dst = (uint4*)dst.ptr; c.total = 1;
	for (int j = 0; j < 4; ++j) {
		const int i = ord[j];
		c.n[j] = (int)(j == 0 ? dst.ne[i] / vl : dst.ne[i]);
		c.ds[j] = j == 0 ? 1 : dst.st[i] / vl; c.ss[j] = j == 0 ? 1 : src.st[i] / vl; c.f[j] = j == 0 ? 1 : f[i];
		c.total *= c.n[j];
	}
	if (!c.total) return true;
	bool pow2 = true;
	for (int j = 1; j < 4; ++j) pow2 = pow2 && (c.f[j] == 1 || c.f[j] == 2 || c.f[j] == 4);
	if (pow2 && c.total < 0xffffffffLL) {
		VecCopy32 q; q.src = c.src; q.dst = c.dst; q.total = (unsigned)c.total;
		for (int j = 0; j < 3; ++j) { q.n[j] = (unsigned)c.n[j]; q.dn[j] = fastdiv32((unsigned)c.n[j]); }
		for (int j = 0; j < 4; ++j) { q.ds[j] = c.ds[j]; q.ss[j] = c.ss[j]; q.fs[j] = c.f[j] == 4 ? 2u : c.f[j] == 2 ? 1u : 0u; }
		// four vectors per thread in flight at most: a grid of a few waves walks the tensor with a grid stride
		const long long blocks = std::min<long long>((c.total + 255) / 256, 148LL * 8 * 4);
		vec_copy32_kernel<<<(unsigned)blocks, 256, 0, s>>>(q);
		g_stats.kernel_launches++;
		return true;
	}
	vec_copy_kernel<<<grid_for(c.total, 256, 2), 256, 0, s>>>(c);
	g_stats.kernel_launches++;
	return true;
}

void k_copy(cudaStream_t s, const View& dst, const View& src)
{
	if (try_vec_copy(s, dst, src)) return;
	Iter4 it = iter_for(dst);
	if (!it.total) return;
	copy_kernel<<<grid_for(it.total, 256), 256, 0, s>>>(v4(dst), v4(src), it);
	g_stats.kernel_launches++;
}

// ------------------------------------------------------------------ binary with broadcast
__global__ void binary_kernel(int op, V4 dst, V4 a, V4 b, Iter4 it)
{
	for (long long lin = blockIdx.x * (long long)blockDim.x + threadIdx.x; lin < it.total;
	     lin += (long long)gridDim.x * blockDim.x) {
		long long i[4], j[4]; decompose(it, lin, i);
		#pragma unroll
		for (int d = 0; d < 4; ++d) j[d] = b.ne[d] == 1 ? 0 : (i[d] % b.ne[d]);
		float x = ldv(a, offs(a, i)), y = ldv(b, offs(b, j));
		stv(dst, offs(dst, i), op == B_MUL ? x * y : x + y);
	}
}
void k_binary(cudaStream_t s, BinOp op, const View& dst, const View& a, const View& b)
{
	Iter4 it = iter_for(dst);
	if (!it.total) return;
	binary_kernel<<<grid_for(it.total, 256), 256, 0, s>>>((int)op, v4(dst), v4(a), v4(b), it);
	g_stats.kernel_launches++;
}

// ------------------------------------------------------------------ unary / scale
__global__ void unary_kernel(int op, float p, V4 dst, V4 src, Iter4 it)
{
	for (long long lin = blockIdx.x * (long long)blockDim.x + threadIdx.x; lin < it.total;
	     lin += (long long)gridDim.x * blockDim.x) {
		long long i[4]; decompose(it, lin, i);
		stv(dst, offs(dst, i), act_apply(op, ldv(src, offs(src, i)), p));
	}
}
void k_unary(cudaStream_t s, UnaryOp op, float param, const View& dst, const View& src)
{
	Iter4 it = iter_for(dst);
	if (!it.total) return;
	unary_kernel<<<grid_for(it.total, 256), 256, 0, s>>>((int)op, param, v4(dst), v4(src), it);
	g_stats.kernel_launches++;
}

// ------------------------------------------------------------------ profiling aid
// One thread spins for a fixed time. The profiled pass (planner.cpp) puts it in front of every timed step so that the
// step's launches are already queued on the device when its start event fires: the event pair then measures the
// kernel, not the host's launch latency (5-10 us for a cluster launch with four tensor maps in its parameters).
__global__ void spin_kernel(long long ns)
{
	unsigned long long t0, t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
	do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); } while ((long long)(t - t0) < ns);
}
void k_spin(cudaStream_t s, int microseconds) { spin_kernel<<<1, 1, 0, s>>>((long long)microseconds * 1000); }

// ------------------------------------------------------------------ nearest upscale (mlblock_nn.c:122)
__global__ void upscale_kernel(V4 dst, V4 src, Iter4 it, int f0, int f1)
{
	for (long long lin = blockIdx.x * (long long)blockDim.x + threadIdx.x; lin < it.total;
	     lin += (long long)gridDim.x * blockDim.x) {
		long long i[4]; decompose(it, lin, i);
		long long j[4] = { i[0] / f0, i[1] / f1, i[2], i[3] };
		stv(dst, offs(dst, i), ldv(src, offs(src, j)));
	}
}
void k_upscale(cudaStream_t s, const View& dst, const View& src)
{
	if (dst.ne[2] == src.ne[2] && dst.ne[3] == src.ne[3] && try_vec_copy(s, dst, src)) return;
	Iter4 it = iter_for(dst);
	upscale_kernel<<<grid_for(it.total, 256), 256, 0, s>>>(v4(dst), v4(src), it,
		(int)(dst.ne[0] / src.ne[0]), (int)(dst.ne[1] / src.ne[1]));
	g_stats.kernel_launches++;
}

// ------------------------------------------------------------------ softmax over dim 0 (generic fallback)
__global__ void softmax_rows_kernel(V4 dst, V4 src, int causal, int n_past)
{
	long long row = blockIdx.x;
	long long i1 = row % src.ne[1], i2 = (row / src.ne[1]) % src.ne[2], i3 = row / (src.ne[1] * src.ne[2]);
	long long so = i1 * src.st[1] + i2 * src.st[2] + i3 * src.st[3];
	long long dof = i1 * dst.st[1] + i2 * dst.st[2] + i3 * dst.st[3];
	long long n = src.ne[0];
	long long lim = causal ? min(n, (long long)n_past + i1 + 1) : n;
	__shared__ float red[32];
	float mx = -INFINITY;
	for (long long i = threadIdx.x; i < lim; i += blockDim.x) mx = fmaxf(mx, ldv(src, so + i * src.st[0]));
	for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(~0u, mx, o));
	if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
	__syncthreads();
	mx = red[0];
	for (int w = 1; w < (blockDim.x >> 5); ++w) mx = fmaxf(mx, red[w]);
	__syncthreads();
	float sum = 0;
	for (long long i = threadIdx.x; i < lim; i += blockDim.x) sum += __expf(ldv(src, so + i * src.st[0]) - mx);
	for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(~0u, sum, o);
	if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
	__syncthreads();
	sum = 0;
	for (int w = 0; w < (blockDim.x >> 5); ++w) sum += red[w];
	float inv = 1.0f / sum;
	for (long long i = threadIdx.x; i < n; i += blockDim.x)
		stv(dst, dof + i * dst.st[0], i < lim ? __expf(ldv(src, so + i * src.st[0]) - mx) * inv : 0.0f);
}
void k_softmax_rows(cudaStream_t s, const View& dst, const View& src, bool causal, int n_past)
{
	long long rows = src.ne[1] * src.ne[2] * src.ne[3];
	softmax_rows_kernel<<<(unsigned)rows, 128, 0, s>>>(v4(dst), v4(src), causal ? 1 : 0, n_past);
	g_stats.kernel_launches++;
}

// ------------------------------------------------------------------ scaled row softmax, f32 scores -> f16 probabilities
// (the wide-head attention path: S = QK^T and O = PV run as tensor-core GEMMs, planner.cpp). One block per row, the row
// lives in registers (NV x 4 floats per thread): one read of the scores, one write of the probabilities.
template <int NV>
__global__ void softmax_f32_f16_kernel(const float* __restrict__ sc, __half* __restrict__ pr, int cols, long long ld_s, long long ld_p, float scale_log2)
{
	const float* row = sc + blockIdx.x * ld_s;
	__half* out = pr + blockIdx.x * ld_p;
	__shared__ float red[32];
	float4 v[NV];
	float mx = -INFINITY;
	#pragma unroll
	for (int i = 0; i < NV; ++i) {
		const int c = (threadIdx.x + i * blockDim.x) * 4;
		if (c < cols) { v[i] = *reinterpret_cast<const float4*>(row + c); mx = fmaxf(mx, fmaxf(fmaxf(v[i].x, v[i].y), fmaxf(v[i].z, v[i].w))); }
	}
	for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(~0u, mx, o));
	if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
	__syncthreads();
	mx = red[0];
	for (int w = 1; w < (int)(blockDim.x >> 5); ++w) mx = fmaxf(mx, red[w]);
	__syncthreads();
	const float m2 = mx * scale_log2;
	float sum = 0.f;
	#pragma unroll
	for (int i = 0; i < NV; ++i) {
		const int c = (threadIdx.x + i * blockDim.x) * 4;
		if (c < cols) {
			v[i].x = exp2f(fmaf(v[i].x, scale_log2, -m2)); v[i].y = exp2f(fmaf(v[i].y, scale_log2, -m2));
			v[i].z = exp2f(fmaf(v[i].z, scale_log2, -m2)); v[i].w = exp2f(fmaf(v[i].w, scale_log2, -m2));
			sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
		}
	}
	for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(~0u, sum, o);
	if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
	__syncthreads();
	sum = 0.f;
	for (int w = 0; w < (int)(blockDim.x >> 5); ++w) sum += red[w];
	const float inv = 1.0f / sum;
	#pragma unroll
	for (int i = 0; i < NV; ++i) {
		const int c = (threadIdx.x + i * blockDim.x) * 4;
		if (c < cols) {
			__half2 a = __floats2half2_rn(v[i].x * inv, v[i].y * inv), b = __floats2half2_rn(v[i].z * inv, v[i].w * inv);
			uint2 o2; o2.x = *reinterpret_cast<uint32_t*>(&a); o2.y = *reinterpret_cast<uint32_t*>(&b);
			*reinterpret_cast<uint2*>(out + c) = o2;
		}
	}
}
bool k_softmax_f32_f16_supported(int64_t cols) { return cols % 4 == 0 && cols <= 256 * 4 * 16; }
void k_softmax_f32_f16(cudaStream_t s, const float* scores, __half* probs, int64_t rows, int64_t cols, int64_t ld_s, int64_t ld_p, float scale)
{
	const float sl2 = scale * 1.4426950408889634f;
	const int nv = (int)((cols / 4 + 255) / 256);
	if (nv <= 1) softmax_f32_f16_kernel<1><<<(unsigned)rows, 256, 0, s>>>(scores, probs, (int)cols, ld_s, ld_p, sl2);
	else if (nv <= 2) softmax_f32_f16_kernel<2><<<(unsigned)rows, 256, 0, s>>>(scores, probs, (int)cols, ld_s, ld_p, sl2);
	else if (nv <= 4) softmax_f32_f16_kernel<4><<<(unsigned)rows, 256, 0, s>>>(scores, probs, (int)cols, ld_s, ld_p, sl2);
	else if (nv <= 8) softmax_f32_f16_kernel<8><<<(unsigned)rows, 256, 0, s>>>(scores, probs, (int)cols, ld_s, ld_p, sl2);
	else softmax_f32_f16_kernel<16><<<(unsigned)rows, 256, 0, s>>>(scores, probs, (int)cols, ld_s, ld_p, sl2);
	g_stats.kernel_launches++;
}

// ------------------------------------------------------------------ get_rows (clip.c:338)
__global__ void get_rows_kernel(V4 dst, V4 table, V4 ids)
{
	long long r = blockIdx.x;  // over dst rows (i1,i2,i3)
	long long i1 = r % dst.ne[1], i2 = (r / dst.ne[1]) % dst.ne[2], i3 = r / (dst.ne[1] * dst.ne[2]);
	int row = ((const int*)ids.ptr)[i1 * ids.st[0] + i2 * ids.st[1] + i3 * ids.st[2]];
	row = max(0, min(row, (int)table.ne[1] - 1));
	for (long long i0 = threadIdx.x; i0 < dst.ne[0]; i0 += blockDim.x)
		stv(dst, i0 * dst.st[0] + i1 * dst.st[1] + i2 * dst.st[2] + i3 * dst.st[3],
			ldv(table, i0 * table.st[0] + row * table.st[1] + i2 * table.st[2]));
}
void k_get_rows(cudaStream_t s, const View& dst, const View& table, const View& ids)
{
	get_rows_kernel<<<(unsigned)(dst.ne[1] * dst.ne[2] * dst.ne[3]), 128, 0, s>>>(v4(dst), v4(table), v4(ids));
	g_stats.kernel_launches++;
}

// ------------------------------------------------------------------ timestep embedding (unet.c:150): cos first
__global__ void timestep_embedding_kernel(V4 dst, V4 ts, int dim, int max_period)
{
	int half = dim / 2;
	long long i = blockIdx.x;
	float t = ldv(ts, i * ts.st[0]);
	for (int j = threadIdx.x; j < half; j += blockDim.x) {
		float freq = expf(-logf((float)max_period) * j / half);
		float arg = t * freq;
		stv(dst, i * dst.st[1] + j * dst.st[0], cosf(arg));
		stv(dst, i * dst.st[1] + (j + half) * dst.st[0], sinf(arg));
	}
	if ((dim & 1) && threadIdx.x == 0) stv(dst, i * dst.st[1] + dim * dst.st[0], 0.0f);
}
void k_timestep_embedding(cudaStream_t s, const View& dst, const View& ts, int dim, int max_period)
{
	timestep_embedding_kernel<<<(unsigned)ts.ne[0], 128, 0, s>>>(v4(dst), v4(ts), dim, max_period);
	g_stats.kernel_launches++;
}

// ------------------------------------------------------------------ generic SIMT GEMM (fallback / debug reference)
// c[m,n,b2,b3] = sum_k a[k,m,b2/r2,b3/r3] * b[k,n,b2,b3]; 32x32 tiles, K step 16.
__global__ void gemm_simt_kernel(V4 c, V4 a, V4 b, int round_b)
{
	__shared__ float sa[16][33], sb[16][33];
	long long M = a.ne[1], N = b.ne[1], K = a.ne[0];
	long long bz = blockIdx.z, i2 = bz % b.ne[2], i3 = bz / b.ne[2];
	long long a2 = i2 / (b.ne[2] / a.ne[2]), a3 = i3 / (b.ne[3] / a.ne[3]);
	long long m0 = blockIdx.x * 32LL, n0 = blockIdx.y * 32LL;
	int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
	float acc[4] = {0, 0, 0, 0};
	for (long long k0 = 0; k0 < K; k0 += 16) {
		for (int e = ty * 32 + tx; e < 16 * 32; e += 256) {
			int kk = e & 15, r = e >> 4;
			long long k = k0 + kk;
			float va = 0, vb = 0;
			if (k < K && m0 + r < M) va = ldv(a, k * a.st[0] + (m0 + r) * a.st[1] + a2 * a.st[2] + a3 * a.st[3]);
			if (k < K && n0 + r < N) {
				vb = ldv(b, k * b.st[0] + (n0 + r) * b.st[1] + i2 * b.st[2] + i3 * b.st[3]);
				if (round_b) vb = __half2float(__float2half_rn(vb));
			}
			sa[kk][r] = va; sb[kk][r] = vb;
		}
		__syncthreads();
		#pragma unroll
		for (int kk = 0; kk < 16; ++kk) {
			float x = sa[kk][tx];
			#pragma unroll
			for (int j = 0; j < 4; ++j) acc[j] += x * sb[kk][ty * 4 + j];
		}
		__syncthreads();
	}
	for (int j = 0; j < 4; ++j) {
		long long m = m0 + tx, n = n0 + ty * 4 + j;
		if (m < M && n < N) stv(c, m * c.st[0] + n * c.st[1] + i2 * c.st[2] + i3 * c.st[3], acc[j]);
	}
}
void k_gemm_simt(cudaStream_t s, const View& c, const View& a, const View& b, bool round_b_f16)
{
	dim3 grid((unsigned)((a.ne[1] + 31) / 32), (unsigned)((b.ne[1] + 31) / 32), (unsigned)(b.ne[2] * b.ne[3]));
	gemm_simt_kernel<<<grid, dim3(32, 8), 0, s>>>(v4(c), v4(a), v4(b), round_b_f16 ? 1 : 0);
	g_stats.kernel_launches++;
}


// ---- deterministic statistics: no floating-point atomics anywhere. Inside a block the per-thread sums are combined in an
// order fixed by indices (shared memory, per channel over the pixel lanes, then per group over the channels). Across blocks
// the per-block group sums are added with INTEGER atomics in fixed point: a double d is split into floor(d) and
// floor((d - floor(d)) * 2^40), both exact 64-bit integers, and integer addition is associative -- the result does not
// depend on the order in which blocks arrive, so a GroupNorm (and with it a whole graph run) is bit-reproducible.
// stats layout: [N][groups][4] 64-bit words = (sum hi, sum lo, sum-of-squares hi, sum-of-squares lo), zeroed before the run.
constexpr double GN_FIX = 1099511627776.0;      // 2^40
// The value added is one block's share of a group (a few thousand f16 values): f32 holds it to 1e-7 relative and the split of an
// f32 is exact (its fraction times 2^40 is an integer). FP64 arithmetic and 64-bit float <-> integer conversions are slow on
// this part (the same split in double cost 12K clk per tile in the GEMM epilogue, profiles/r2_ncu_gemm.md).
__device__ __forceinline__ void gn_fix_add(unsigned long long* dst, float v)
{
	const float hi = floorf(v);
	const long long ihi = (long long)hi, ilo = (long long)((v - hi) * 1099511627776.0f);
	atomicAdd(dst, (unsigned long long)ihi);        // two's complement: negative sums wrap correctly
	atomicAdd(dst + 1, (unsigned long long)ilo);
}
__device__ __forceinline__ double gn_fix_get(const unsigned long long* src)
{ return (double)(long long)src[0] + (double)(long long)src[1] * (1.0 / GN_FIX); }

__device__ __forceinline__ void gn_block_finish(const float* su, const float* sq, bool active, int ch_local, int plane, int planes, int slab, int slab_chunks,
	int chunks, int cpg, int groups, int n, unsigned long long* stats, float* smf)
{
	// smf: [planes][slab_chunks * 8][2] floats, then chs[slab_chunks * 8][2] floats: gn_stats_smem() bytes
	const int nch = slab_chunks * 8;
	float* chs = smf + (size_t)planes * nch * 2;                 // [nch][2] per-channel sums of this block
	if (plane < planes) {
		#pragma unroll
		for (int j = 0; j < 8; ++j) {
			float* d = smf + ((size_t)plane * nch + ch_local * 8 + j) * 2;
			d[0] = active ? su[j] : 0.f; d[1] = active ? sq[j] : 0.f;
		}
	}
	__syncthreads();
	for (int c = threadIdx.x; c < nch; c += blockDim.x) {
		float a = 0.f, b = 0.f;
		for (int pl = 0; pl < planes; ++pl) { const float* d = smf + ((size_t)pl * nch + c) * 2; a += d[0]; b += d[1]; }
		chs[c * 2] = a; chs[c * 2 + 1] = b;
	}
	__syncthreads();
	const int c_lo = slab * nch, c_hi = min(chunks * 8, c_lo + nch);        // channels of this slab
	for (int g = threadIdx.x; g < groups; g += blockDim.x) {
		const int g_lo = max(g * cpg, c_lo), g_hi = min((g + 1) * cpg, c_hi);
		if (g_lo >= g_hi) continue;
		float a = 0.f, b = 0.f;
		for (int c = g_lo; c < g_hi; ++c) { a += chs[(c - c_lo) * 2]; b += chs[(c - c_lo) * 2 + 1]; }
		unsigned long long* dst = stats + ((size_t)n * groups + g) * 4;
		gn_fix_add(dst, a); gn_fix_add(dst + 2, b);
	}
}

// ------------------------------------------------------------------ GroupNorm (+affine, +SiLU), channels-last f16
// (mlblock_nn.c:78-103 + ggml_silu_inplace :136,147). Reference statistics are double sums; here
// per-thread f32 partials over <= 64 elements are combined in double.
// Pass 1: block = 256 threads covers a tile of pixels x all channels (coalesced), per-group partial
// sums in shared memory, one double atomicAdd pair per group per block.
// Pass 1. Grid: (pixel tiles x channel slabs, images). A thread owns ONE 8-channel chunk and walks the
// pixels of the tile with a fixed stride, so its partial sums stay in registers; they are flushed
// once per thread (run-length combined per group) to shared memory, then once per block to the
// double-precision statistics in global memory.
template <typename T>
__global__ void gn_stats_kernel(const T* __restrict__ x, long long HW, int C, int cpg, int groups,
	long long img_stride, long long pix_stride, unsigned long long* __restrict__ stats, int pix_per_block, int slab_chunks, int nslabs)
{
	extern __shared__ __align__(16) float sm[];
	const int n = blockIdx.y;
	const int tile = blockIdx.x / nslabs, slab = blockIdx.x - tile * nslabs;
	const int chunks = C / 8;
	const int ch = slab * slab_chunks + (int)(threadIdx.x % slab_chunks);
	const int plane = threadIdx.x / slab_chunks, planes = blockDim.x / slab_chunks;
	const long long p0 = (long long)tile * pix_per_block, p1 = min(HW, p0 + pix_per_block);
	float su[8], sq[8];
	#pragma unroll
	for (int j = 0; j < 8; ++j) { su[j] = 0.f; sq[j] = 0.f; }
	const bool active = plane < planes && ch < chunks;
	if (active) {
		const T* base = x + n * img_stride + ch * 8;
		for (long long p = p0 + plane; p < p1; p += planes) {
			const T* ptr = base + p * pix_stride;
			float v[8];
			if (sizeof(T) == 2) {
				uint4 raw = *reinterpret_cast<const uint4*>(ptr);
				const __half2* h = reinterpret_cast<const __half2*>(&raw);
				#pragma unroll
				for (int j = 0; j < 4; ++j) { float2 f = __half22float2(h[j]); v[2*j] = f.x; v[2*j+1] = f.y; }
			} else {
				#pragma unroll
				for (int j = 0; j < 8; ++j) v[j] = (float)ptr[j];
			}
			#pragma unroll
			for (int j = 0; j < 8; ++j) { su[j] += v[j]; sq[j] += v[j] * v[j]; }
		}
	}
	gn_block_finish(su, sq, active, (int)(threadIdx.x % slab_chunks), plane, planes, slab, slab_chunks, chunks, cpg, groups, n, stats, sm);
}

template <typename TI, typename TO>
__global__ void gn_apply_kernel(const TI* __restrict__ x, TO* __restrict__ y, long long HW, int C, int cpg, int groups,
	long long img_stride, long long pix_stride, long long oimg_stride, long long opix_stride,
	const float* __restrict__ gamma, const float* __restrict__ beta, const unsigned long long* __restrict__ stats,
	float eps, int silu)
{
	extern __shared__ float sm[];  // mean[groups], rstd[groups]
	int n = blockIdx.y;
	double cnt = (double)HW * cpg;
	for (int g = threadIdx.x; g < groups; g += blockDim.x) {
		double su = gn_fix_get(stats + ((long long)n * groups + g) * 4), sq = gn_fix_get(stats + ((long long)n * groups + g) * 4 + 2);
		double mean = su / cnt, var = sq / cnt - mean * mean;
		if (var < 0) var = 0;
		sm[g] = (float)mean;
		sm[groups + g] = (float)(1.0 / sqrt(var + (double)eps));
	}
	__syncthreads();
	int chunks = C / 8;
	long long total = HW * chunks;
	for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
		long long p = e / chunks; int ch = (int)(e - p * chunks);
		const TI* ptr = x + n * img_stride + p * pix_stride + ch * 8;
		float v[8];
		if (sizeof(TI) == 2) {
			uint4 raw = *reinterpret_cast<const uint4*>(ptr);
			const __half2* h = reinterpret_cast<const __half2*>(&raw);
			#pragma unroll
			for (int j = 0; j < 4; ++j) { float2 f = __half22float2(h[j]); v[2*j] = f.x; v[2*j+1] = f.y; }
		} else {
			#pragma unroll
			for (int j = 0; j < 8; ++j) v[j] = (float)ptr[j];
		}
		#pragma unroll
		for (int j = 0; j < 8; ++j) {
			int c = ch * 8 + j, g = c / cpg;
			float t = (v[j] - sm[g]) * sm[groups + g];
			if (gamma) t = t * gamma[c] + (beta ? beta[c] : 0.f);
			if (silu) t = t / (1.0f + __expf(-t));
			v[j] = t;
		}
		TO* optr = y + n * oimg_stride + p * opix_stride + ch * 8;
		if (sizeof(TO) == 2) {
			uint4 raw; __half2* h = reinterpret_cast<__half2*>(&raw);
			#pragma unroll
			for (int j = 0; j < 4; ++j) h[j] = __floats2half2_rn(v[2*j], v[2*j+1]);
			*reinterpret_cast<uint4*>(optr) = raw;
		} else {
			#pragma unroll
			for (int j = 0; j < 8; ++j) optr[j] = (TO)v[j];
		}
	}
}

// ---- fast path (f16 in, f16 out): a thread owns ONE 8-channel chunk for its whole life, so mean/rstd/gamma/beta
// fold into 8 scale + 8 shift registers once and the pixel loop is load -> 8 FMA (-> SiLU) -> store, four pixels
// in flight per thread. Same (chunk, plane) thread mapping as the statistics pass: a warp touches contiguous memory.
__device__ __forceinline__ void h8_to_f(const uint4& raw, float* v)
{
	const __half2* h = reinterpret_cast<const __half2*>(&raw);
	#pragma unroll
	for (int j = 0; j < 4; ++j) { float2 f = __half22float2(h[j]); v[2*j] = f.x; v[2*j+1] = f.y; }
}
__device__ __forceinline__ uint4 f_to_h8(const float* v)
{
	uint4 raw; __half2* h = reinterpret_cast<__half2*>(&raw);
	#pragma unroll
	for (int j = 0; j < 4; ++j) h[j] = __floats2half2_rn(v[2*j], v[2*j+1]);
	return raw;
}

constexpr int GN_BLOCKS_PER_SM = 5;      // both fast kernels are compiled for 5 resident blocks of <= 256 threads per SM (48 registers)

__global__ void __launch_bounds__(256, GN_BLOCKS_PER_SM)
gn_stats_fast_kernel(const __half* __restrict__ x, long long HW, int C, int cpg, int groups,
	long long img_stride, long long pix_stride, unsigned long long* __restrict__ stats, int pix_per_block, int slab_chunks, int nslabs)
{
	extern __shared__ __align__(16) float sm[];
	const int n = blockIdx.y;
	const int tile = blockIdx.x / nslabs, slab = blockIdx.x - tile * nslabs;
	const int chunks = C / 8;
	const int ch = slab * slab_chunks + (int)(threadIdx.x % slab_chunks);
	const int plane = threadIdx.x / slab_chunks, planes = blockDim.x / slab_chunks;
	const long long p0 = (long long)tile * pix_per_block, p1 = min(HW, p0 + pix_per_block);
	float su[8], sq[8];
	#pragma unroll
	for (int j = 0; j < 8; ++j) { su[j] = 0.f; sq[j] = 0.f; }
	const bool active = ch < chunks;
	if (active) {
		const __half* base = x + n * img_stride + ch * 8;
		long long p = p0 + plane;
		for (; p + 3 * planes < p1; p += 4 * planes) {
			uint4 r[4];
			#pragma unroll
			for (int u = 0; u < 4; ++u) r[u] = *reinterpret_cast<const uint4*>(base + (p + u * planes) * pix_stride);
			#pragma unroll
			for (int u = 0; u < 4; ++u) { float v[8]; h8_to_f(r[u], v);
				#pragma unroll
				for (int j = 0; j < 8; ++j) { su[j] += v[j]; sq[j] = fmaf(v[j], v[j], sq[j]); } }
		}
		for (; p < p1; p += planes) {
			float v[8]; h8_to_f(*reinterpret_cast<const uint4*>(base + p * pix_stride), v);
			#pragma unroll
			for (int j = 0; j < 8; ++j) { su[j] += v[j]; sq[j] = fmaf(v[j], v[j], sq[j]); }
		}
	}
	gn_block_finish(su, sq, active, (int)(threadIdx.x % slab_chunks), plane, planes, slab, slab_chunks, chunks, cpg, groups, n, stats, sm);
}

// x * sigmoid(x) = x * (0.5 + 0.5 tanh(x / 2)): one MUFU op per element (tanh.approx) instead of ex2 + rcp
__device__ __forceinline__ float silu_tanh(float t)
{
	float th; asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(0.5f * t));
	const float h = 0.5f * t;
	return fmaf(h, th, h);
}

template <bool SILU>
__global__ void __launch_bounds__(256, GN_BLOCKS_PER_SM)
gn_apply_fast_kernel(const __half* __restrict__ x, __half* __restrict__ y, long long HW, int C, int cpg, int groups,
	long long img_stride, long long pix_stride, long long oimg_stride, long long opix_stride,
	const float* __restrict__ gamma, const float* __restrict__ beta, const unsigned long long* __restrict__ stats,
	float eps, int pix_per_block, int slab_chunks, int nslabs)
{
	extern __shared__ float sm[];  // mean[groups], rstd[groups]: the double-precision finish runs once per group and block
	const int n = blockIdx.y;
	const int tile = blockIdx.x / nslabs, slab = blockIdx.x - tile * nslabs;
	const double cnt = (double)HW * cpg;
	for (int g = threadIdx.x; g < groups; g += blockDim.x) {
		const double su = gn_fix_get(stats + ((long long)n * groups + g) * 4), sq = gn_fix_get(stats + ((long long)n * groups + g) * 4 + 2);
		const double mean = su / cnt; double var = sq / cnt - mean * mean; if (var < 0) var = 0;
		sm[g] = (float)mean;
		sm[groups + g] = (float)(1.0 / sqrt(var + (double)eps));
	}
	__syncthreads();
	const int chunks = C / 8;
	const int ch = slab * slab_chunks + (int)(threadIdx.x % slab_chunks);
	const int plane = threadIdx.x / slab_chunks, planes = blockDim.x / slab_chunks;
	if (ch >= chunks) return;
	// scale / shift of my 8 channels: one division (the channels are consecutive: the group index only steps), vector loads of
	// the affine parameters -- this set-up was 23 % of the kernel's samples with 8 divisions and 16 scalar loads per thread
	float sc[8], sh[8], ga[8], be[8];
	#pragma unroll
	for (int j = 0; j < 8; ++j) { ga[j] = 1.f; be[j] = 0.f; }
	if (gamma) {
		if (!((uintptr_t)gamma & 15)) {
			const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + ch * 8)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + ch * 8 + 4));
			ga[0] = g0.x; ga[1] = g0.y; ga[2] = g0.z; ga[3] = g0.w; ga[4] = g1.x; ga[5] = g1.y; ga[6] = g1.z; ga[7] = g1.w;
		} else {
			#pragma unroll
			for (int j = 0; j < 8; ++j) ga[j] = gamma[ch * 8 + j];
		}
		if (beta) {
			if (!((uintptr_t)beta & 15)) {
				const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + ch * 8)), b1 = __ldg(reinterpret_cast<const float4*>(beta + ch * 8 + 4));
				be[0] = b0.x; be[1] = b0.y; be[2] = b0.z; be[3] = b0.w; be[4] = b1.x; be[5] = b1.y; be[6] = b1.z; be[7] = b1.w;
			} else {
				#pragma unroll
				for (int j = 0; j < 8; ++j) be[j] = beta[ch * 8 + j];
			}
		}
	}
	int g = (ch * 8) / cpg, rem = ch * 8 - g * cpg;              // channel ch * 8 + j sits in group g at offset rem
	#pragma unroll
	for (int j = 0; j < 8; ++j) {
		sc[j] = sm[groups + g] * ga[j]; sh[j] = be[j] - sm[g] * sc[j];
		if (++rem == cpg) { rem = 0; ++g; }
	}
	const long long p0 = (long long)tile * pix_per_block, p1 = min(HW, p0 + pix_per_block);
	const __half* base = x + n * img_stride + ch * 8;
	__half* obase = y + n * oimg_stride + ch * 8;
	auto apply = [&](const uint4& raw) {
		float v[8]; h8_to_f(raw, v);
		#pragma unroll
		for (int j = 0; j < 8; ++j) { float t = fmaf(v[j], sc[j], sh[j]); if (SILU) t = silu_tanh(t); v[j] = t; }
		return f_to_h8(v);
	};
	long long p = p0 + plane;
	for (; p + 3 * planes < p1; p += 4 * planes) {
		uint4 r[4];
		#pragma unroll
		for (int u = 0; u < 4; ++u) r[u] = *reinterpret_cast<const uint4*>(base + (p + u * planes) * pix_stride);
		#pragma unroll
		for (int u = 0; u < 4; ++u) *reinterpret_cast<uint4*>(obase + (p + u * planes) * opix_stride) = apply(r[u]);
	}
	for (; p < p1; p += planes)
		*reinterpret_cast<uint4*>(obase + p * opix_stride) = apply(*reinterpret_cast<const uint4*>(base + p * pix_stride));
}


// ---- small tensors (f16, channels-last): ONE kernel, one block per (image, group). The block's slice (HW pixels x cpg
// channels, <= GN_SMALL_MAXV half2 per thread) is read once into registers, reduced exactly in two passes (mean, then
// centred sum of squares, f32 per thread, combined in double across the block) and written back normalised: no
// statistics buffer, no atomics, no second read. Used for the smallest levels (8x8, 16x16x640 per image), where the
// two-kernel path is launch-latency bound.
constexpr int GN_SMALL_MAXV = 64;
template <bool SILU, int NV>     // NV: half2 values per thread (upper bound, compile time so that they stay in registers)
__global__ void __launch_bounds__(256)
gn_small_kernel(const __half* __restrict__ x, __half* __restrict__ y, int HW, int cpg, int groups,
	long long img_stride, long long pix_stride, long long oimg_stride, long long opix_stride,
	const float* __restrict__ gamma, const float* __restrict__ beta, float eps, unsigned c2n_mul)
{
	__shared__ float red[8];        // f32 throughout: the second pass is centred (no cancellation), and FP64 division / square root per thread cost more than the slice
	const int g = blockIdx.x, n = blockIdx.y, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
	const int c2n = cpg >> 1, total = HW * c2n;                 // half2 units per pixel of this group / in the slice
	const __half* xb = x + n * img_stride + (long long)g * cpg;
	__half* yb = y + n * oimg_stride + (long long)g * cpg;
	// element e = tid + 256 i of the slice is (pixel p = e / c2n, channel pair c2): one multiply-high instead of a division
	__half2 v[NV];
	float sum = 0.f;
	#pragma unroll
	for (int i = 0; i < NV; ++i) {
		const int e = tid + i * 256;
		if (e < total) {
			const int p = c2n_mul ? (int)__umulhi((unsigned)e, c2n_mul) : e, c2 = e - p * c2n;
			v[i] = *reinterpret_cast<const __half2*>(xb + p * pix_stride + 2 * c2);
			const float2 f = __half22float2(v[i]);
			sum += f.x + f.y;
		}
	}
	auto block_sum = [&](float s) -> float {
		#pragma unroll
		for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(~0u, s, o);
		__syncthreads();                                         // red[] of the previous reduction has been consumed
		if (lane == 0) red[w] = s;
		__syncthreads();
		float t = 0.f;
		#pragma unroll
		for (int i = 0; i < 8; ++i) t += red[i];
		return t;
	};
	const float inv_cnt = 1.0f / ((float)HW * (float)cpg);
	const float mean = block_sum(sum) * inv_cnt;
	float sq = 0.f;
	#pragma unroll
	for (int i = 0; i < NV; ++i)
		if (tid + i * 256 < total) { const float2 f = __half22float2(v[i]); const float a = f.x - mean, b = f.y - mean; sq = fmaf(a, a, sq); sq = fmaf(b, b, sq); }
	const float rstd = rsqrtf(block_sum(sq) * inv_cnt + eps);
	#pragma unroll
	for (int i = 0; i < NV; ++i) {
		const int e = tid + i * 256;
		if (e < total) {
			const int p = c2n_mul ? (int)__umulhi((unsigned)e, c2n_mul) : e, c2 = e - p * c2n, c = g * cpg + 2 * c2;
			const float g0 = gamma ? gamma[c] : 1.f, g1 = gamma ? gamma[c + 1] : 1.f;
			const float b0 = (gamma && beta) ? beta[c] : 0.f, b1 = (gamma && beta) ? beta[c + 1] : 0.f;
			const float2 f = __half22float2(v[i]);
			float t0 = fmaf((f.x - mean) * rstd, g0, b0), t1 = fmaf((f.y - mean) * rstd, g1, b1);
			if (SILU) { t0 = silu_tanh(t0); t1 = silu_tanh(t1); }
			*reinterpret_cast<__half2*>(yb + p * opix_stride + 2 * c2) = __floats2half2_rn(t0, t1);
		}
	}
}

template <int NV>
static void gn_small_launch(cudaStream_t s, bool silu, dim3 grid, const __half* x, __half* y, int HW, int cpg, int groups,
	long long is, long long ps, long long ois, long long ops, const float* gamma, const float* beta, float eps)
{
	const unsigned c2n = (unsigned)(cpg / 2), mul = c2n <= 1 ? 0u : (unsigned)((0x100000000ull + c2n - 1) / c2n);   // exact while e * c2n < 2^32
	if (silu) gn_small_kernel<true, NV><<<grid, 256, 0, s>>>(x, y, HW, cpg, groups, is, ps, ois, ops, gamma, beta, eps, mul);
	else gn_small_kernel<false, NV><<<grid, 256, 0, s>>>(x, y, HW, cpg, groups, is, ps, ois, ops, gamma, beta, eps, mul);
}

static size_t gn_stats_smem(int planes, int slab_chunks, int groups)
{ (void)groups; return (size_t)planes * slab_chunks * 8 * 2 * sizeof(float) + (size_t)slab_chunks * 8 * 2 * sizeof(float); }

// launch geometry of the statistics pass (shared with the planner's scratch sizing)
struct GnGeom { int threads, planes, slab_chunks, nslabs, pix_per_block; unsigned grid_x; };
static GnGeom gn_geom(long long HW, int C, long long N, bool fast)
{
	GnGeom g;
	const int chunks = C / 8;
	if (fast) {
		g.nslabs = (chunks + 255) / 256; g.slab_chunks = (chunks + g.nslabs - 1) / g.nslabs;
		g.planes = std::max(1, 256 / g.slab_chunks); g.threads = g.slab_chunks * g.planes;
		// One exact wave: 148 SMs x GN_BLOCKS_PER_SM resident blocks share the pixels evenly (no tail wave, every SM equally
		// loaded); small tensors get one pixel row of `planes` pixels per block at least.
		const long long target = 148LL * GN_BLOCKS_PER_SM;
		const long long tiles_want = std::max<long long>(1, target / std::max<long long>(1, N * g.nslabs));
		long long ppb = (HW + tiles_want - 1) / tiles_want;
		ppb = std::max<long long>(g.planes, (ppb + g.planes - 1) / g.planes * g.planes);
		g.pix_per_block = (int)std::min<long long>(ppb, 1 << 30);
	} else {
		g.threads = 256;
		g.slab_chunks = std::min(chunks, 64); g.nslabs = (chunks + g.slab_chunks - 1) / g.slab_chunks;
		g.planes = g.threads / g.slab_chunks;
		g.pix_per_block = g.planes * 16;                  // 16 pixels per thread
		if (HW * N * g.nslabs < 148LL * 4 * g.pix_per_block) g.pix_per_block = g.planes * 4;   // small tensors: more, smaller blocks
	}
	g.grid_x = (unsigned)(((HW + g.pix_per_block - 1) / g.pix_per_block) * g.nslabs);
	return g;
}

// stats: [N][groups][4] 64-bit words (fixed-point sums, see gn_fix_add), zeroed before the run
void k_groupnorm(cudaStream_t s, const View& dst, const View& src, const float* gamma, const float* beta,
	int groups, float eps, bool silu, unsigned long long* stats, bool stats_ready)
{
	int C = (int)src.ne[2]; long long W = src.ne[0], H = src.ne[1], N = src.ne[3], HW = W * H;
	int cpg = (C + groups - 1) / groups;
	if (src.st[2] != 1 || dst.st[2] != 1 || C % 8 || src.st[1] != W * src.st[0] || dst.st[1] != W * dst.st[0])
		B200_FATAL("k_groupnorm: unsupported layout (C=%d)", C);
	if (src.dt == DT_F16 && dst.dt == DT_F16 && !((uintptr_t)src.ptr & 15) && !((uintptr_t)dst.ptr & 15) &&
		src.st[0] % 8 == 0 && dst.st[0] % 8 == 0 && src.st[3] % 8 == 0 && dst.st[3] % 8 == 0) {
		// small slices: single pass, one block per (image, group)
		static const bool small_on = !(getenv("GGML_B200_GN_SMALL") && atoi(getenv("GGML_B200_GN_SMALL")) == 0);
		const long long slice2 = HW * (cpg / 2);                 // half2 units per (image, group)
		if (!stats_ready && small_on && C == groups * cpg && cpg % 2 == 0 && slice2 <= 256LL * 12 && HW < (1 << 24)) {          // measured: beyond ~12 values per thread the two-kernel path wins
			dim3 grid((unsigned)groups, (unsigned)N);
			const __half* xp = (const __half*)src.ptr; __half* yp = (__half*)dst.ptr;
			const int nv = (int)((slice2 + 255) / 256);
			if (nv <= 8) gn_small_launch<8>(s, silu, grid, xp, yp, (int)HW, cpg, groups, src.st[3], src.st[0], dst.st[3], dst.st[0], gamma, beta, eps);
			else if (nv <= 16) gn_small_launch<16>(s, silu, grid, xp, yp, (int)HW, cpg, groups, src.st[3], src.st[0], dst.st[3], dst.st[0], gamma, beta, eps);
			else if (nv <= 32) gn_small_launch<32>(s, silu, grid, xp, yp, (int)HW, cpg, groups, src.st[3], src.st[0], dst.st[3], dst.st[0], gamma, beta, eps);
			else gn_small_launch<GN_SMALL_MAXV>(s, silu, grid, xp, yp, (int)HW, cpg, groups, src.st[3], src.st[0], dst.st[3], dst.st[0], gamma, beta, eps);
			g_stats.kernel_launches++;
			return;
		}
		const GnGeom G = gn_geom(HW, C, N, true);
		const int slab_chunks = G.slab_chunks, nslabs = G.nslabs, threads = G.threads, pix_per_block = G.pix_per_block;
		dim3 grid(G.grid_x, (unsigned)N);
		const size_t smem = groups * 2 * sizeof(float);
		static bool attr = false;
		if (!attr) { CUDA_CHECK(cudaFuncSetAttribute(gn_stats_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024)); attr = true; }
		if (!stats_ready)
			gn_stats_fast_kernel<<<grid, threads, gn_stats_smem(G.planes, slab_chunks, groups), s>>>((const __half*)src.ptr, HW, C, cpg, groups, src.st[3], src.st[0], stats, pix_per_block, slab_chunks, nslabs);
		if (silu)
			gn_apply_fast_kernel<true><<<grid, threads, smem, s>>>((const __half*)src.ptr, (__half*)dst.ptr, HW, C, cpg, groups, src.st[3], src.st[0],
				dst.st[3], dst.st[0], gamma, beta, stats, eps, pix_per_block, slab_chunks, nslabs);
		else
			gn_apply_fast_kernel<false><<<grid, threads, smem, s>>>((const __half*)src.ptr, (__half*)dst.ptr, HW, C, cpg, groups, src.st[3], src.st[0],
				dst.st[3], dst.st[0], gamma, beta, stats, eps, pix_per_block, slab_chunks, nslabs);
		g_stats.kernel_launches += stats_ready ? 1 : 2;
		return;
	}
	if (stats_ready) B200_FATAL("k_groupnorm: epilogue statistics are only consumed by the f16 path");
	const GnGeom G = gn_geom(HW, C, N, false);
	const int threads = G.threads, slab_chunks = G.slab_chunks, nslabs = G.nslabs, pix_per_block = G.pix_per_block;
	dim3 g1(G.grid_x, (unsigned)N);
	size_t smem = groups * 2 * sizeof(float);
	const size_t smem1 = gn_stats_smem(G.planes, slab_chunks, groups);
	if (src.dt == DT_F16)
		gn_stats_kernel<__half><<<g1, threads, smem1, s>>>((const __half*)src.ptr, HW, C, cpg, groups, src.st[3], src.st[0], stats, pix_per_block, slab_chunks, nslabs);
	else
		gn_stats_kernel<float><<<g1, threads, smem1, s>>>((const float*)src.ptr, HW, C, cpg, groups, src.st[3], src.st[0], stats, pix_per_block, slab_chunks, nslabs);
	long long total = HW * (C / 8);
	dim3 g2((unsigned)std::min<long long>((total + threads - 1) / threads, 148 * 8), (unsigned)N);
	if (src.dt == DT_F16 && dst.dt == DT_F16)
		gn_apply_kernel<__half, __half><<<g2, threads, smem, s>>>((const __half*)src.ptr, (__half*)dst.ptr, HW, C, cpg, groups,
			src.st[3], src.st[0], dst.st[3], dst.st[0], gamma, beta, stats, eps, silu ? 1 : 0);
	else if (src.dt == DT_F32 && dst.dt == DT_F16)
		gn_apply_kernel<float, __half><<<g2, threads, smem, s>>>((const float*)src.ptr, (__half*)dst.ptr, HW, C, cpg, groups,
			src.st[3], src.st[0], dst.st[3], dst.st[0], gamma, beta, stats, eps, silu ? 1 : 0);
	else if (src.dt == DT_F16 && dst.dt == DT_F32)
		gn_apply_kernel<__half, float><<<g2, threads, smem, s>>>((const __half*)src.ptr, (float*)dst.ptr, HW, C, cpg, groups,
			src.st[3], src.st[0], dst.st[3], dst.st[0], gamma, beta, stats, eps, silu ? 1 : 0);
	else
		gn_apply_kernel<float, float><<<g2, threads, smem, s>>>((const float*)src.ptr, (float*)dst.ptr, HW, C, cpg, groups,
			src.st[3], src.st[0], dst.st[3], dst.st[0], gamma, beta, stats, eps, silu ? 1 : 0);
	g_stats.kernel_launches += 2;
}

// ------------------------------------------------------------------ LayerNorm (+affine), one warp per row
// (mlblock_nn.c:58-75; eps 1e-5). Row cached in registers: two exact passes, f32.
// NCH = 8-element chunks per lane (row length <= 256 * NCH); 16-byte loads/stores for f16 rows.
template <typename TI, typename TO, int NCH>
__global__ void layernorm_kernel(const TI* __restrict__ x, TO* __restrict__ y, long long rows, int C,
	long long ld_in, long long ld_out, const float* __restrict__ gamma, const float* __restrict__ beta, float eps)
{
	long long row = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
	if (row >= rows) return;
	const int lane = threadIdx.x & 31, chunks = C >> 3;
	const TI* xr = x + row * ld_in;
	float v[NCH][8];
	float sum = 0.f;
	#pragma unroll
	for (int i = 0; i < NCH; ++i) {
		const int ch = lane + i * 32;
		if (ch < chunks) {
			if (sizeof(TI) == 2) {
				uint4 raw = *reinterpret_cast<const uint4*>(xr + ch * 8);
				const __half2* h = reinterpret_cast<const __half2*>(&raw);
				#pragma unroll
				for (int j = 0; j < 4; ++j) { float2 f = __half22float2(h[j]); v[i][2*j] = f.x; v[i][2*j+1] = f.y; }
			} else {
				#pragma unroll
				for (int j = 0; j < 8; ++j) v[i][j] = (float)xr[ch * 8 + j];
			}
			#pragma unroll
			for (int j = 0; j < 8; ++j) sum += v[i][j];
		}
	}
	for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(~0u, sum, o);
	const float mean = sum / C;
	float sq = 0.f;
	#pragma unroll
	for (int i = 0; i < NCH; ++i)
		if (lane + i * 32 < chunks) {
			#pragma unroll
			for (int j = 0; j < 8; ++j) { v[i][j] -= mean; sq += v[i][j] * v[i][j]; }
		}
	for (int o = 16; o; o >>= 1) sq += __shfl_xor_sync(~0u, sq, o);
	const float rstd = rsqrtf(sq / C + eps);
	TO* yr = y + row * ld_out;
	#pragma unroll
	for (int i = 0; i < NCH; ++i) {
		const int ch = lane + i * 32;
		if (ch < chunks) {
			float t[8];
			#pragma unroll
			for (int j = 0; j < 8; ++j) {
				t[j] = v[i][j] * rstd;
				if (gamma) t[j] = t[j] * __ldg(gamma + ch * 8 + j) + (beta ? __ldg(beta + ch * 8 + j) : 0.f);
			}
			if (sizeof(TO) == 2) {
				uint4 raw; __half2* h = reinterpret_cast<__half2*>(&raw);
				#pragma unroll
				for (int j = 0; j < 4; ++j) h[j] = __floats2half2_rn(t[2*j], t[2*j+1]);
				*reinterpret_cast<uint4*>(yr + ch * 8) = raw;
			} else {
				#pragma unroll
				for (int j = 0; j < 8; ++j) yr[ch * 8 + j] = (TO)t[j];
			}
		}
	}
}

// ---- fast path (f16 -> f16): persistent grid (exactly the resident blocks of the chip). A row is owned by LPR lanes
// (LPR * NCH 8-element chunks, chunk = lane_in_row + i * LPR), so a warp works on 32 / LPR rows at once with every lane
// busy: the row widths of the diffusion models are 320 * 2^k = 40 * 2^k chunks, which a lane-per-chunk mapping would
// leave 37 % idle. gamma / beta of the lane's chunks live in registers for the whole kernel; the next row group is in
// flight while the current one is reduced (two exact passes, f32, butterfly over the LPR lanes) and stored.
template <int LPR, int NCH>
__global__ void __launch_bounds__(256)
layernorm_fast_kernel(const __half* __restrict__ x, __half* __restrict__ y, long long rows, int C,
	long long ld_in, long long ld_out, const float* __restrict__ gamma, const float* __restrict__ beta, float eps)
{
	constexpr int RW = 32 / LPR;                        // rows per warp and iteration
	const int lane = threadIdx.x & 31, chunks = C >> 3;
	const int li = lane % LPR, sub = lane / LPR;
	const long long wstride = (long long)gridDim.x * (blockDim.x >> 5) * RW;
	long long row = (blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5)) * RW + sub;
	uint4 cur[NCH], nxt[NCH];
	auto load = [&](long long r, uint4* dst) {
		if (r < rows) {
			#pragma unroll
			for (int i = 0; i < NCH; ++i) { const int ch = li + i * LPR; if (ch < chunks) dst[i] = *reinterpret_cast<const uint4*>(x + r * ld_in + ch * 8); }
		}
	};
	load(row, cur);
	float ga[NCH][8], be[NCH][8];
	#pragma unroll
	for (int i = 0; i < NCH; ++i) {
		const int ch = li + i * LPR;
		#pragma unroll
		for (int j = 0; j < 8; ++j) { ga[i][j] = 1.f; be[i][j] = 0.f; }
		if (ch < chunks && gamma) {
			const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + ch * 8)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + ch * 8 + 4));
			ga[i][0] = g0.x; ga[i][1] = g0.y; ga[i][2] = g0.z; ga[i][3] = g0.w; ga[i][4] = g1.x; ga[i][5] = g1.y; ga[i][6] = g1.z; ga[i][7] = g1.w;
			if (beta) {
				const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + ch * 8)), b1 = __ldg(reinterpret_cast<const float4*>(beta + ch * 8 + 4));
				be[i][0] = b0.x; be[i][1] = b0.y; be[i][2] = b0.z; be[i][3] = b0.w; be[i][4] = b1.x; be[i][5] = b1.y; be[i][6] = b1.z; be[i][7] = b1.w;
			}
		}
	}
	const float inv_c = 1.0f / C;
	// every lane of the warp runs the same number of iterations (the shuffles need all of them): loop on the group base
	for (long long base = row - sub; base < rows; base += wstride, row += wstride) {
		load(row + wstride, nxt);
		float v[NCH][8];
		float sum = 0.f;
		#pragma unroll
		for (int i = 0; i < NCH; ++i) {
			if (li + i * LPR < chunks && row < rows) { h8_to_f(cur[i], v[i]);
				#pragma unroll
				for (int j = 0; j < 8; ++j) sum += v[i][j]; }
			else {
				#pragma unroll
				for (int j = 0; j < 8; ++j) v[i][j] = 0.f;
			}
		}
		#pragma unroll
		for (int o = LPR / 2; o; o >>= 1) sum += __shfl_xor_sync(~0u, sum, o);
		const float mean = sum * inv_c;
		float sq = 0.f;
		#pragma unroll
		for (int i = 0; i < NCH; ++i)
			if (li + i * LPR < chunks) {
				#pragma unroll
				for (int j = 0; j < 8; ++j) { v[i][j] -= mean; sq = fmaf(v[i][j], v[i][j], sq); }
			}
		#pragma unroll
		for (int o = LPR / 2; o; o >>= 1) sq += __shfl_xor_sync(~0u, sq, o);
		const float rstd = rsqrtf(sq * inv_c + eps);
		if (row < rows) {
			#pragma unroll
			for (int i = 0; i < NCH; ++i) {
				const int ch = li + i * LPR;
				if (ch < chunks) {
					float t[8];
					#pragma unroll
					for (int j = 0; j < 8; ++j) t[j] = fmaf(v[i][j] * rstd, ga[i][j], be[i][j]);
					*reinterpret_cast<uint4*>(y + row * ld_out + ch * 8) = f_to_h8(t);
				}
			}
		}
		#pragma unroll
		for (int i = 0; i < NCH; ++i) cur[i] = nxt[i];
	}
}

template <int LPR, int NCH>
static void layernorm_fast_launch(cudaStream_t s, const __half* x, __half* y, long long rows, int C, long long ldi, long long ldo,
	const float* g, const float* b, float eps)
{
	static int bps = 0;       // resident blocks per SM of this instantiation
	if (!bps) {
		if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, layernorm_fast_kernel<LPR, NCH>, 256, 0) != cudaSuccess || bps < 1) { cudaGetLastError(); bps = 1; }
	}
	constexpr int RW = 32 / LPR;
	const long long groups = (rows + RW - 1) / RW;
	const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((groups + 7) / 8, 148LL * bps));
	layernorm_fast_kernel<LPR, NCH><<<grid, 256, 0, s>>>(x, y, rows, C, ldi, ldo, g, b, eps);
}

// lanes per row and chunks per lane for a row of `chunks` 8-element chunks: the fewest chunks per lane (<= 5) with no idle lane
static bool layernorm_fast_dispatch(cudaStream_t s, const __half* x, __half* y, long long rows, int C, long long ldi, long long ldo,
	const float* g, const float* b, float eps)
{
	const int chunks = C / 8;
	#define LN_TRY(LPR, NCH) if (chunks <= LPR * NCH && chunks > LPR * (NCH - 1)) { layernorm_fast_launch<LPR, NCH>(s, x, y, rows, C, ldi, ldo, g, b, eps); return true; }
	static const int mode = getenv("GGML_B200_LN_MODE") ? atoi(getenv("GGML_B200_LN_MODE")) : 0;   // 1: 3 chunks per lane (fewer registers, some idle lanes)
	if (mode == 1 && chunks % 5 == 0 && chunks / 5 <= 16) { switch (chunks / 5) { case 8: LN_TRY(16, 3) break; case 16: LN_TRY(32, 3) break; } }
	if (chunks % 5 == 0 && chunks / 5 <= 32 && ((chunks / 5) & (chunks / 5 - 1)) == 0) {      // 40 * 2^k chunks: 320, 640, 1280 (and 40, 80, 160)
		switch (chunks / 5) { case 1: LN_TRY(1, 5) break; case 2: LN_TRY(2, 5) break; case 4: LN_TRY(4, 5) break; case 8: LN_TRY(8, 5) break;
			case 16: LN_TRY(16, 5) break; case 32: LN_TRY(32, 5) break; }
	}
	if (chunks % 3 == 0 && chunks / 3 <= 32 && ((chunks / 3) & (chunks / 3 - 1)) == 0) {      // 96 chunks: 768 (CLIP ViT-L)
		switch (chunks / 3) { case 8: LN_TRY(8, 3) break; case 16: LN_TRY(16, 3) break; case 32: LN_TRY(32, 3) break; }
	}
	if ((chunks & (chunks - 1)) == 0) {                                                        // powers of two: 1024 (OpenCLIP ViT-H), 2048
		switch (chunks) { case 8: LN_TRY(8, 1) break; case 16: LN_TRY(16, 1) break; case 32: LN_TRY(32, 1) break; case 64: LN_TRY(32, 2) break; case 128: LN_TRY(32, 4) break; }
	}
	LN_TRY(32, 1) LN_TRY(32, 2) LN_TRY(32, 3) LN_TRY(32, 4) LN_TRY(32, 5)
	#undef LN_TRY
	return false;
}

// scalar fallback for row lengths that are not multiples of 8 (or unaligned rows)
template <typename TI, typename TO>
__global__ void layernorm_scalar_kernel(const TI* __restrict__ x, TO* __restrict__ y, long long rows, int C,
	long long ld_in, long long ld_out, const float* __restrict__ gamma, const float* __restrict__ beta, float eps)
{
	long long row = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
	if (row >= rows) return;
	const int lane = threadIdx.x & 31;
	const TI* xr = x + row * ld_in;
	float sum = 0.f;
	for (int c = lane; c < C; c += 32) sum += (float)xr[c];
	for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(~0u, sum, o);
	const float mean = sum / C;
	float sq = 0.f;
	for (int c = lane; c < C; c += 32) { float d = (float)xr[c] - mean; sq += d * d; }
	for (int o = 16; o; o >>= 1) sq += __shfl_xor_sync(~0u, sq, o);
	const float rstd = rsqrtf(sq / C + eps);
	TO* yr = y + row * ld_out;
	for (int c = lane; c < C; c += 32) {
		float t = ((float)xr[c] - mean) * rstd;
		if (gamma) t = t * gamma[c] + (beta ? beta[c] : 0.f);
		yr[c] = (TO)t;
	}
}

template <typename TI, typename TO>
static void layernorm_launch(cudaStream_t s, const void* x, void* y, long long rows, int C, long long ldi, long long ldo,
	const float* g, const float* b, float eps)
{
	int wpb = 8;
	unsigned grid = (unsigned)((rows + wpb - 1) / wpb);
	bool vec = C % 8 == 0 && ldi % 8 == 0 && ldo % 8 == 0 && ((uintptr_t)x % 16) == 0 && ((uintptr_t)y % 16) == 0 && C <= 2048;
	if (!vec) { layernorm_scalar_kernel<TI, TO><<<grid, wpb * 32, 0, s>>>((const TI*)x, (TO*)y, rows, C, ldi, ldo, g, b, eps); return; }
	if (sizeof(TI) == 2 && sizeof(TO) == 2 && C <= 256 * 5 && (!g || !((uintptr_t)g & 15)) && (!b || !((uintptr_t)b & 15))) {
		if (layernorm_fast_dispatch(s, (const __half*)x, (__half*)y, rows, C, ldi, ldo, g, b, eps)) return;
	}
	if (C <= 256 * 2) layernorm_kernel<TI, TO, 2><<<grid, wpb * 32, 0, s>>>((const TI*)x, (TO*)y, rows, C, ldi, ldo, g, b, eps);
	else if (C <= 256 * 5) layernorm_kernel<TI, TO, 5><<<grid, wpb * 32, 0, s>>>((const TI*)x, (TO*)y, rows, C, ldi, ldo, g, b, eps);
	else layernorm_kernel<TI, TO, 8><<<grid, wpb * 32, 0, s>>>((const TI*)x, (TO*)y, rows, C, ldi, ldo, g, b, eps);
}

void k_layernorm(cudaStream_t s, const View& dst, const View& src, const float* gamma, const float* beta, float eps)
{
	int C = (int)src.ne[0];
	// rows must have a uniform pitch
	long long rows = src.ne[1] * src.ne[2] * src.ne[3];
	auto uniform = [](const View& v) {
		long long p = v.st[1];
		if (v.ne[1] == 1) p = v.ne[2] > 1 ? v.st[2] : v.st[3];
		bool ok = v.st[0] == 1;
		if (v.ne[2] > 1 && v.ne[1] > 1) ok = ok && v.st[2] == v.st[1] * v.ne[1];
		if (v.ne[3] > 1) ok = ok && v.st[3] == (v.ne[2] > 1 ? v.st[2] * v.ne[2] : v.st[1] * v.ne[1]);
		return ok ? p : -1;
	};
	long long ldi = rows > 1 ? uniform(src) : C, ldo = rows > 1 ? uniform(dst) : C;
	if (ldi < 0 || ldo < 0) B200_FATAL("k_layernorm: rows are not uniformly strided");
	if (src.dt == DT_F16 && dst.dt == DT_F16) layernorm_launch<__half, __half>(s, src.ptr, dst.ptr, rows, C, ldi, ldo, gamma, beta, eps);
	else if (src.dt == DT_F32 && dst.dt == DT_F16) layernorm_launch<float, __half>(s, src.ptr, dst.ptr, rows, C, ldi, ldo, gamma, beta, eps);
	else if (src.dt == DT_F16 && dst.dt == DT_F32) layernorm_launch<__half, float>(s, src.ptr, dst.ptr, rows, C, ldi, ldo, gamma, beta, eps);
	else layernorm_launch<float, float>(s, src.ptr, dst.ptr, rows, C, ldi, ldo, gamma, beta, eps);
	g_stats.kernel_launches++;
}

// ------------------------------------------------------------------ GEGLU gate (mlblock_nn.c:159-172)
// h: rows of 2d f16 (value | gate), dst rows of d f16. 8 elements per thread.
__global__ void geglu_kernel(const __half* __restrict__ h, __half* __restrict__ y, long long rows, int d,
	long long ldh, long long ldy)
{
	int chunks = d / 8;
	long long total = rows * chunks;
	for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
		long long r = e / chunks; int ch = (int)(e - r * chunks);
		uint4 xv = *reinterpret_cast<const uint4*>(h + r * ldh + ch * 8);
		uint4 gv = *reinterpret_cast<const uint4*>(h + r * ldh + d + ch * 8);
		const __half2* xh = reinterpret_cast<const __half2*>(&xv);
		const __half2* gh = reinterpret_cast<const __half2*>(&gv);
		uint4 ov; __half2* oh = reinterpret_cast<__half2*>(&ov);
		#pragma unroll
		for (int j = 0; j < 4; ++j) {
			float2 x = __half22float2(xh[j]), g = __half22float2(gh[j]);
			oh[j] = __floats2half2_rn(x.x * act_apply(U_GELU, g.x, 0.f), x.y * act_apply(U_GELU, g.y, 0.f));
		}
		*reinterpret_cast<uint4*>(y + r * ldy + ch * 8) = ov;
	}
}
void k_geglu(cudaStream_t s, const View& dst, const View& h)
{
	int d = (int)dst.ne[0];
	long long rows = dst.ne[1] * dst.ne[2] * dst.ne[3];
	if (h.dt != DT_F16 || dst.dt != DT_F16 || d % 8 || h.st[0] != 1 || dst.st[0] != 1)
		B200_FATAL("k_geglu: unsupported layout");
	long long total = rows * (d / 8);
	geglu_kernel<<<grid_for(total, 256), 256, 0, s>>>((const __half*)h.ptr, (__half*)dst.ptr, rows, d, h.st[1], dst.st[1]);
	g_stats.kernel_launches++;
}

// ------------------------------------------------------------------ GEGLU weight / bias row permutation (once per weight version)
__global__ void geglu_rows_prep_kernel(V4 dst, V4 src, long long D)
{
	const long long K = src.ne[0], total = 2 * D * K;
	for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
		const long long r = e / K, k = e - r * K;
		const long long blk = r >> 5, i = r & 31;
		const long long sr = i < 16 ? blk * 16 + i : D + blk * 16 + (i - 16);
		stv(dst, r * K + k, ldv(src, k * src.st[0] + sr * src.st[1]));
	}
}
void k_geglu_rows_prep(cudaStream_t s, void* dst, const View& src, int64_t D)
{
	// src: [K, 2D] (weights) or [2D] viewed as [1, 2D] (bias)
	View sv = src;
	if (src.ne[1] == 1 && src.ne[0] == 2 * D) { sv.ne[0] = 1; sv.st[0] = 1; sv.ne[1] = 2 * D; sv.st[1] = src.st[0]; }
	if (sv.ne[1] != 2 * D) B200_FATAL("k_geglu_rows_prep: expected %lld rows, got %lld", (long long)(2 * D), (long long)sv.ne[1]);
	View dv = sv; dv.ptr = dst; dv.st[0] = 1; dv.st[1] = sv.ne[0];
	long long total = 2 * D * sv.ne[0];
	geglu_rows_prep_kernel<<<grid_for(total, 256, 2), 256, 0, s>>>(v4(dv), v4(sv), D);
	g_stats.kernel_launches++;
}

// ------------------------------------------------------------------ im2col for strided / narrow convs
__global__ void im2col_kernel(__half* __restrict__ col, long long kpad, V4 x, int KW, int KH,
	int s0, int s1, int p0, int p1, int d0, int d1, long long OW, long long OH)
{
	long long C = x.ne[2], K = (long long)KW * KH * C;
	long long M = OW * OH * x.ne[3];
	long long total = M * kpad;
	for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
		long long m = e / kpad, k = e - m * kpad;
		float v = 0.f;
		if (k < K) {
			long long c = k % C, tap = k / C; int kw = (int)(tap % KW), kh = (int)(tap / KW);
			long long ow = m % OW, oh = (m / OW) % OH, n = m / (OW * OH);
			long long iw = ow * s0 + kw * d0 - p0, ih = oh * s1 + kh * d1 - p1;
			if (iw >= 0 && iw < x.ne[0] && ih >= 0 && ih < x.ne[1])
				v = ldv(x, iw * x.st[0] + ih * x.st[1] + c * x.st[2] + n * x.st[3]);
		}
		col[e] = __float2half_rn(v);
	}
}
// channels-last f16 input with Cin % 8 == 0: every (output pixel, tap) is a contiguous run of Cin channels -> 16-byte copies
__global__ void im2col_vec_kernel(uint4* __restrict__ col, long long kpad8, const __half* __restrict__ x, long long W, long long H, long long C8,
	long long sw, long long sh, long long sn, int KW, int KH, int s0, int s1, int p0, int p1, int d0, int d1, long long OW, long long OH, long long M)
{
	const long long K8 = (long long)KW * KH * C8, total = M * kpad8;
	for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
		const long long m = e / kpad8, k = e - m * kpad8;
		uint4 v = make_uint4(0u, 0u, 0u, 0u);
		if (k < K8) {
			const long long tap = k / C8, c8 = k - tap * C8; const int kw = (int)(tap % KW), kh = (int)(tap / KW);
			const long long ow = m % OW, oh = (m / OW) % OH, n = m / (OW * OH);
			const long long iw = ow * s0 + kw * d0 - p0, ih = oh * s1 + kh * d1 - p1;
			if (iw >= 0 && iw < W && ih >= 0 && ih < H) v = __ldg(reinterpret_cast<const uint4*>(x + iw * sw + ih * sh + n * sn) + c8);
		}
		col[e] = v;
	}
}

void k_im2col(cudaStream_t s, __half* col, int64_t kpad, const View& x, int KW, int KH,
	int s0, int s1, int p0, int p1, int d0, int d1, int64_t OW, int64_t OH)
{
	if (x.dt == DT_F16 && x.st[2] == 1 && x.ne[2] % 8 == 0 && kpad % 8 == 0 && x.st[0] % 8 == 0 && x.st[1] % 8 == 0 && (x.ne[3] == 1 || x.st[3] % 8 == 0) &&
		!((uintptr_t)x.ptr & 15) && !((uintptr_t)col & 15)) {
		const long long M = OW * OH * x.ne[3], total = M * (kpad / 8);
		im2col_vec_kernel<<<grid_for(total, 256, 2), 256, 0, s>>>((uint4*)col, kpad / 8, (const __half*)x.ptr, x.ne[0], x.ne[1], x.ne[2] / 8,
			x.st[0], x.st[1], x.st[3], KW, KH, s0, s1, p0, p1, d0, d1, OW, OH, M);
		g_stats.kernel_launches++;
		return;
	}
	long long total = OW * OH * x.ne[3] * kpad;
	im2col_kernel<<<grid_for(total, 256, 4), 256, 0, s>>>(col, kpad, v4(x), KW, KH, s0, s1, p0, p1, d0, d1, OW, OH);
	g_stats.kernel_launches++;
}

// conv weight ggml [KW,KH,Cin,Cout] -> rows [Cout][(kh*KW+kw)*Cin + c] with pitch kpad (zero padded)
__global__ void conv_weight_prep_kernel(__half* __restrict__ dst, long long kpad, V4 w)
{
	long long KW = w.ne[0], KH = w.ne[1], C = w.ne[2], OC = w.ne[3], K = KW * KH * C;
	long long total = OC * kpad;
	for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
		long long oc = e / kpad, k = e - oc * kpad;
		float v = 0.f;
		if (k < K) {
			long long c = k % C, tap = k / C, kw = tap % KW, kh = tap / KW;
			v = ldv(w, kw * w.st[0] + kh * w.st[1] + c * w.st[2] + oc * w.st[3]);
		}
		dst[e] = __float2half_rn(v);
	}
}
void k_conv_weight_prep(cudaStream_t s, __half* dst, int64_t kpad, const View& w)
{
	long long total = w.ne[3] * kpad;
	conv_weight_prep_kernel<<<grid_for(total, 256, 2), 256, 0, s>>>(dst, kpad, v4(w));
	g_stats.kernel_launches++;
}

// ------------------------------------------------------------------ 3x3 convolution with 3 or 4 output channels (direct, SIMT)
// conv_out of the UNet (320 -> 4, unet.c:338) and the last convolution of the VAE decoder (128 -> 3, vae.c:262) are the two
// launches the tensor-core path cannot serve: a 128 x 16 MMA tile wastes 3/4 of its columns and the 4- / 3-channel output row
// is too narrow for the TMA store, so they ran on the non-persistent kernel at 7-18 TFLOP/s (81 us and 1035 us). Here a block
// owns 32 x 8 output pixels (one per thread): the input halo (34 x 10 pixels) goes through shared memory in chunks of 64
// channels (pixel pitch 144 B: the 16-byte loads of a quarter-warp fall on 32 different banks), the weights of the chunk sit
// beside it in f32 and are read as warp-wide broadcasts, the products accumulate in f32 pairs (FFMA2).
__device__ __forceinline__ unsigned long long f2_pack(float a, float b) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ unsigned long long f2_fma(unsigned long long a, unsigned long long b, unsigned long long c)
{ unsigned long long r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
constexpr int CS_TW = 32, CS_TH = 8, CS_CC = 64, CS_PITCH = CS_CC * 2 + 16, CS_HALO = (CS_TW + 2) * (CS_TH + 2);
template <int COUT>
__global__ void __launch_bounds__(256, 3)
conv3x3_small_kernel(const __half* __restrict__ x, const __half* __restrict__ w, const float* __restrict__ bias, __half* __restrict__ y,
	int H, int W, int Cin, int tiles_w, int tiles_h)
{
	extern __shared__ __align__(16) uint8_t cs_sm[];
	uint8_t* sx = cs_sm;                                            // [CS_HALO pixels][CS_PITCH bytes]
	float* sw = reinterpret_cast<float*>(cs_sm + CS_HALO * CS_PITCH);   // [COUT][9][CS_CC]
	const int tid = threadIdx.x, px = tid & 31, py = tid >> 5;
	int b = blockIdx.x;
	const int tx = b % tiles_w; b /= tiles_w;
	const int ty = b % tiles_h; const int n = b / tiles_h;
	const int x0 = tx * CS_TW, y0 = ty * CS_TH;
	const __half* xin = x + (long long)n * H * W * Cin;
	unsigned long long acc[COUT];
	#pragma unroll
	for (int co = 0; co < COUT; ++co) acc[co] = 0ull;
	for (int c0 = 0; c0 < Cin; c0 += CS_CC) {
		__syncthreads();                                            // the previous chunk has been consumed
		for (int idx = tid; idx < CS_HALO * 8; idx += 256) {
			const int pix = idx >> 3, part = idx & 7;
			const int gy = y0 - 1 + pix / (CS_TW + 2), gx = x0 - 1 + pix % (CS_TW + 2);
			uint4 v = make_uint4(0u, 0u, 0u, 0u);                   // zero padding outside the image
			if (gy >= 0 && gy < H && gx >= 0 && gx < W) v = __ldg(reinterpret_cast<const uint4*>(xin + ((long long)gy * W + gx) * Cin + c0 + part * 8));
			*reinterpret_cast<uint4*>(sx + pix * CS_PITCH + part * 16) = v;
		}
		for (int idx = tid; idx < COUT * 9 * CS_CC; idx += 256) {
			const int co = idx / (9 * CS_CC), r = idx - co * 9 * CS_CC, tap = r / CS_CC, c = r - tap * CS_CC;
			sw[idx] = __half2float(w[(long long)co * 9 * Cin + (long long)tap * Cin + c0 + c]);
		}
		__syncthreads();
		#pragma unroll
		for (int tap = 0; tap < 9; ++tap) {
			const uint8_t* p = sx + ((py + tap / 3) * (CS_TW + 2) + px + tap % 3) * CS_PITCH;
			#pragma unroll 2
			for (int part = 0; part < 8; ++part) {
				const uint4 raw = *reinterpret_cast<const uint4*>(p + part * 16);
				const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
				unsigned long long f[4];
				#pragma unroll
				for (int j = 0; j < 4; ++j) { const float2 t = __half22float2(h2[j]); f[j] = f2_pack(t.x, t.y); }
				#pragma unroll
				for (int co = 0; co < COUT; ++co) {
					const float4* wp = reinterpret_cast<const float4*>(sw + (co * 9 + tap) * CS_CC + part * 8);
					const float4 w0 = wp[0], w1 = wp[1];
					acc[co] = f2_fma(f[0], f2_pack(w0.x, w0.y), acc[co]);
					acc[co] = f2_fma(f[1], f2_pack(w0.z, w0.w), acc[co]);
					acc[co] = f2_fma(f[2], f2_pack(w1.x, w1.y), acc[co]);
					acc[co] = f2_fma(f[3], f2_pack(w1.z, w1.w), acc[co]);
				}
			}
		}
	}
	const int gx = x0 + px, gy = y0 + py;
	if (gx < W && gy < H) {
		__half* o = y + (((long long)n * H + gy) * W + gx) * COUT;
		float r[COUT];
		#pragma unroll
		for (int co = 0; co < COUT; ++co) {
			float a, c; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(c) : "l"(acc[co]));
			r[co] = a + c + (bias ? bias[co] : 0.f);
		}
		if (COUT == 4) {
			uint2 pk; __half2* hp = reinterpret_cast<__half2*>(&pk);
			hp[0] = __floats2half2_rn(r[0], r[1]); hp[1] = __floats2half2_rn(r[2], r[3 % COUT]);
			*reinterpret_cast<uint2*>(o) = pk;
		} else {
			#pragma unroll
			for (int co = 0; co < COUT; ++co) o[co] = __float2half_rn(r[co]);
		}
	}
}

bool k_conv3x3_small_supported(int64_t Cin, int64_t Cout)
{
	// opt-in (GGML_B200_CONV_SMALL=1): parity-green on hardware (tests/test_ops_gpu.py::test_conv3x3_small_direct, 3e-4 vs the oracle)
	// but measured at 89 us on the UNet's conv_out (64 x 64 x 320 -> 4, 16 latents: 256 blocks of 8 warps leave the chip latency-bound),
	// no better than the tensor-core fallback there (the VAE's last convolution, 512 x 512 x 128 -> 3 x 8 images, runs at 623 us against
	// 1021 us); it needs smaller pixel tiles and a double-buffered halo before it becomes the default
	const char* e = getenv("GGML_B200_CONV_SMALL");              // read per plan (a test switches it inside one process)
	const bool on = e && atoi(e) != 0;
	return on && (Cout == 3 || Cout == 4) && Cin % CS_CC == 0;
}
// x: [N][H][W][Cin] f16 dense, w: [Cout][9 * Cin] f16 (kh, kw, Cin order: k_conv_weight_prep), y: [N][H][W][Cout] f16 dense
void k_conv3x3_small(cudaStream_t s, __half* y, const __half* x, const __half* w, const float* bias, int64_t N, int64_t H, int64_t W, int64_t Cin, int64_t Cout)
{
	const int tiles_w = (int)((W + CS_TW - 1) / CS_TW), tiles_h = (int)((H + CS_TH - 1) / CS_TH);
	const unsigned grid = (unsigned)(N * tiles_h * tiles_w);
	const size_t smem = (size_t)CS_HALO * CS_PITCH + (size_t)Cout * 9 * CS_CC * sizeof(float);
	static bool attr = false;
	if (!attr) {
		CUDA_CHECK(cudaFuncSetAttribute(conv3x3_small_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
		CUDA_CHECK(cudaFuncSetAttribute(conv3x3_small_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
		attr = true;
	}
	if (Cout == 4) conv3x3_small_kernel<4><<<grid, 256, smem, s>>>(x, w, bias, y, (int)H, (int)W, (int)Cin, tiles_w, tiles_h);
	else conv3x3_small_kernel<3><<<grid, 256, smem, s>>>(x, w, bias, y, (int)H, (int)W, (int)Cin, tiles_w, tiles_h);
	g_stats.kernel_launches++;
}

}  // namespace b200
