// attention.cu -- fused scaled-dot-product attention (no score materialisation).
// Replaces the reference's unfused ggml_nn_attention (ggml_extend.c:200-221): QK^T, scale,
// optional causal mask, softmax, PV -- three extra passes over an [nk,nq,heads] f32 tensor there,
// one kernel with online softmax here.
//
// k_attention_simt: generic streaming kernel (any head dim <= 512, any strides, f16/f32 operands),
// warp-per-query online softmax with K/V tiles staged in shared memory. It is the correctness
// baseline and the path for shapes the tcgen05 kernel does not cover.
#include "kernels.h"

namespace b200 {

struct AV { const void* ptr; int dt; long long st_d, st_t, st_h, st_b; };

__device__ __forceinline__ float av_ld(const AV& a, long long off)
{
	return a.dt == DT_F16 ? __half2float(((const __half*)a.ptr)[off]) : ((const float*)a.ptr)[off];
}

constexpr int ATT_QPB = 16;   // queries per block (4 per warp)
constexpr int ATT_KT = 32;    // keys per tile

template <int DPL>  // output dims per lane = ceil(d/32)
__global__ void __launch_bounds__(128) attention_simt_kernel(AV q, AV k, AV v, void* o, int o_dt,
	long long so_d, long long so_t, long long so_h, long long so_b,
	int d, long long nq, long long nk, int H, float scale, int causal)
{
	extern __shared__ __align__(16) unsigned char smraw[];
	const int kp = d + 2;                                  // padded K row (halves): odd word pitch
	__half* sK = reinterpret_cast<__half*>(smraw);         // [KT][kp]
	__half* sV = sK + ATT_KT * kp;                         // [KT][d]
	float*  sQ = reinterpret_cast<float*>(sV + ATT_KT * d + ((ATT_KT * (kp + d)) & 1)); // [QPB][d]

	const int h = blockIdx.y % H, b = blockIdx.y / H;
	const long long q0 = (long long)blockIdx.x * ATT_QPB;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

	for (int e = threadIdx.x; e < ATT_QPB * d; e += blockDim.x) {
		int qi = e / d, dd = e - qi * d;
		long long t = q0 + qi;
		sQ[e] = t < nq ? av_ld(q, dd * q.st_d + t * q.st_t + h * q.st_h + b * q.st_b) * scale : 0.f;
	}

	float m[4], l[4], acc[4][DPL];
	#pragma unroll
	for (int i = 0; i < 4; ++i) { m[i] = -INFINITY; l[i] = 0.f;
		#pragma unroll
		for (int j = 0; j < DPL; ++j) acc[i][j] = 0.f; }

	long long k_end = nk;
	if (causal) k_end = min(nk, q0 + ATT_QPB);  // keys beyond the last query of the block are masked
	for (long long k0 = 0; k0 < k_end; k0 += ATT_KT) {
		__syncthreads();
		for (int e = threadIdx.x; e < ATT_KT * d; e += blockDim.x) {
			int j = e / d, dd = e - j * d;
			long long t = k0 + j;
			float kv = 0.f, vv = 0.f;
			if (t < nk) {
				kv = av_ld(k, dd * k.st_d + t * k.st_t + h * k.st_h + b * k.st_b);
				vv = av_ld(v, dd * v.st_d + t * v.st_t + h * v.st_h + b * v.st_b);
			}
			sK[j * kp + dd] = __float2half_rn(kv);
			sV[j * d + dd] = __float2half_rn(vv);
		}
		__syncthreads();
		#pragma unroll
		for (int qi = 0; qi < 4; ++qi) {
			long long t = q0 + warp * 4 + qi;
			const float* qr = sQ + (warp * 4 + qi) * d;
			float s = 0.f;
			const __half* kr = sK + lane * kp;
			for (int dd = 0; dd < d; dd += 2) {
				float2 kk = __half22float2(*reinterpret_cast<const __half2*>(kr + dd));
				s += qr[dd] * kk.x + qr[dd + 1] * kk.y;
			}
			long long kt = k0 + lane;
			bool valid = kt < nk && (!causal || kt <= t);
			if (!valid) s = -INFINITY;
			float mx = s;
			for (int o2 = 16; o2; o2 >>= 1) mx = fmaxf(mx, __shfl_xor_sync(~0u, mx, o2));
			float mnew = fmaxf(m[qi], mx);
			float corr = (m[qi] == -INFINITY) ? 0.f : __expf(m[qi] - mnew);
			float p = (mnew == -INFINITY || !valid) ? 0.f : __expf(s - mnew);
			float ps = p;
			for (int o2 = 16; o2; o2 >>= 1) ps += __shfl_xor_sync(~0u, ps, o2);
			l[qi] = l[qi] * corr + ps;
			m[qi] = mnew;
			#pragma unroll
			for (int j = 0; j < DPL; ++j) acc[qi][j] *= corr;
			for (int j = 0; j < ATT_KT; ++j) {
				float pj = __shfl_sync(~0u, p, j);
				#pragma unroll
				for (int i = 0; i < DPL; ++i) {
					int dd = lane + 32 * i;
					if (dd < d) acc[qi][i] += pj * __half2float(sV[j * d + dd]);
				}
			}
		}
	}
	#pragma unroll
	for (int qi = 0; qi < 4; ++qi) {
		long long t = q0 + warp * 4 + qi;
		if (t >= nq) continue;
		float inv = l[qi] > 0.f ? 1.0f / l[qi] : 0.f;
		#pragma unroll
		for (int i = 0; i < DPL; ++i) {
			int dd = lane + 32 * i;
			if (dd < d) {
				long long off = dd * so_d + t * so_t + h * so_h + b * so_b;
				float r = acc[qi][i] * inv;
				if (o_dt == DT_F16) ((__half*)o)[off] = __float2half_rn(r); else ((float*)o)[off] = r;
			}
		}
	}
}

static void attention_simt(cudaStream_t s, const View& o, const View& q, const View& k, const View& v, float scale, bool causal)
{
	int d = (int)q.ne[0]; long long nq = q.ne[1], nk = k.ne[1]; int H = (int)q.ne[2], B = (int)q.ne[3];
	if (d > 512 || (d & 1)) B200_FATAL("attention: head dim %d unsupported", d);
	AV aq { q.ptr, q.dt, q.st[0], q.st[1], q.st[2], q.st[3] };
	AV ak { k.ptr, k.dt, k.st[0], k.st[1], k.st[2], k.st[3] };
	// v arrives as the ggml V^T view [nk, d, H, B]
	AV av { v.ptr, v.dt, v.st[1], v.st[0], v.st[2], v.st[3] };
	size_t smem = (size_t)ATT_KT * (d + 2) * 2 + (size_t)ATT_KT * d * 2 + 8 + (size_t)ATT_QPB * d * 4;
	dim3 grid((unsigned)((nq + ATT_QPB - 1) / ATT_QPB), (unsigned)(H * B));
	int dpl = (d + 31) / 32;
#define ATT_LAUNCH(DPL) do { \
	CUDA_CHECK(cudaFuncSetAttribute(attention_simt_kernel<DPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
	attention_simt_kernel<DPL><<<grid, 128, smem, s>>>(aq, ak, av, o.ptr, (int)o.dt, o.st[0], o.st[1], o.st[2], o.st[3], \
		d, nq, nk, H, scale, causal ? 1 : 0); } while (0)
	if (dpl <= 2) ATT_LAUNCH(2);
	else if (dpl <= 3) ATT_LAUNCH(3);
	else if (dpl <= 5) ATT_LAUNCH(5);
	else ATT_LAUNCH(16);
#undef ATT_LAUNCH
	g_stats.kernel_launches++;
}

void k_attention(cudaStream_t s, const View& o, const View& q, const View& k, const View& v, float scale, bool causal)
{
	attention_simt(s, o, q, k, v, scale, causal);
}

}  // namespace b200
