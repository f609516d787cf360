// tools/pipe_rates.cu -- development aid: issue cost (cycles per element, one or two warps per SM sub-partition)
// of the instruction mixes the attention softmax can be built from. Run on the GPU box.
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2_mufu(float x) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x)); return x; }
// exp2 on the FMA pipe (Cody-Waite split + degree-3 polynomial), x <= 0
__device__ __forceinline__ float ex2_poly(float x)
{
	x = fmaxf(x, -126.0f);
	const float magic = 12582912.0f;              // 1.5 * 2^23: rounds to nearest integer in the low mantissa bits
	float xr = x + magic;
	float n = xr - magic;
	float f = x - n;                              // in [-0.5, 0.5]
	float p = fmaf(f, 0.0555041f, 0.2402265f);
	p = fmaf(p, f, 0.6931472f);
	p = fmaf(p, f, 1.0f);
	return __int_as_float(__float_as_int(p) + (__float_as_int(xr) << 23));
}
// packed path: scale in f32, pack to f16x2, ONE ex2.approx.f16x2 per pair (result is already the f16 probability pair)
template <int MODE> __global__ void kh(float* out, long long* cyc, float seed)
{
	float x[32]; unsigned h[16];
	for (int i = 0; i < 32; ++i) x[i] = seed + i * 0.001f + threadIdx.x * 1e-5f;
	for (int i = 0; i < 16; ++i) h[i] = 0;
	unsigned acch = 0; float acc = 0.f;
	__syncthreads();
	long long t0 = clock64();
	#pragma unroll 1
	for (int it = 0; it < 64; ++it) {
		#pragma unroll
		for (int i = 0; i < 32; i += 2) {
			float a = fmaf(x[i], 0.999f, seed), b = fmaf(x[i + 1], 0.999f, seed);
			unsigned pk, e;
			asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(pk) : "f"(b), "f"(a));
			asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(e) : "r"(pk));
			if (MODE == 20) asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(acch) : "r"(e));
			if (MODE == 22) { float lo, hi; asm volatile("{.reg .b16 l, h; mov.b32 {l, h}, %2; cvt.f32.f16 %0, l; cvt.f32.f16 %1, h;}" : "=f"(lo), "=f"(hi) : "r"(e)); acc += lo + hi; }
			h[i >> 1] ^= e;
			x[i] = -a; x[i + 1] = -b;
		}
	}
	long long t1 = clock64();
	for (int i = 0; i < 32; ++i) acc += x[i];
	for (int i = 0; i < 16; ++i) acc += (float)h[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = acc + (float)acch;
	if (threadIdx.x == 0 && blockIdx.x == 0) cyc[MODE - 16] = t1 - t0;
}
template <int MODE> __global__ void k(float* out, long long* cyc, float seed)
{
	float x[32]; unsigned h[16];
	for (int i = 0; i < 32; ++i) x[i] = seed + i * 0.001f + threadIdx.x * 1e-5f;
	for (int i = 0; i < 16; ++i) h[i] = 0;
	float acc = 0.f;
	__syncthreads();
	long long t0 = clock64();
	#pragma unroll 1
	for (int it = 0; it < 64; ++it) {
		#pragma unroll
		for (int i = 0; i < 32; i += 2) {
			float a = fmaf(x[i], 0.999f, seed), b = fmaf(x[i + 1], 0.999f, seed);
			constexpr int NPOLY = MODE == 7 ? 8 : MODE == 8 ? 2 : MODE == 9 ? 3 : MODE == 10 ? 4 : 0;   // of every 8 elements
			if ((i & 7) < NPOLY) a = ex2_poly(a); else a = ex2_mufu(a);
			if (((i + 1) & 7) < NPOLY) b = ex2_poly(b); else b = ex2_mufu(b);
			if (MODE != 5) acc += a + b;
			unsigned pk = 0;
			if (MODE != 6) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(pk) : "f"(a), "f"(b));
			h[i >> 1] ^= pk;
			x[i] = -a; x[i + 1] = -b;
		}
	}
	long long t1 = clock64();
	for (int i = 0; i < 32; ++i) acc += x[i];
	for (int i = 0; i < 16; ++i) acc += (float)h[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
	if (threadIdx.x == 0 && blockIdx.x == 0) cyc[MODE] = t1 - t0;
}
int main()
{
	float* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 512);
	for (int w = 1; w <= 2; ++w) {
		int threads = 128 * w;
		k<4><<<148, threads>>>(out, cyc, -1.f); k<5><<<148, threads>>>(out, cyc, -1.f); k<6><<<148, threads>>>(out, cyc, -1.f);
		k<7><<<148, threads>>>(out, cyc, -1.f); k<8><<<148, threads>>>(out, cyc, -1.f); k<9><<<148, threads>>>(out, cyc, -1.f); k<10><<<148, threads>>>(out, cyc, -1.f);
		kh<20><<<148, threads>>>(out, cyc + 16, -1.f); kh<21><<<148, threads>>>(out, cyc + 16, -1.f); kh<22><<<148, threads>>>(out, cyc + 16, -1.f);
		long long h[16]; cudaMemcpy(h, cyc, 128, cudaMemcpyDeviceToHost);
		long long hh[8]; cudaMemcpy(hh, cyc + 16, 64, cudaMemcpyDeviceToHost);
		auto c = [&](int m) { return h[m] / (64.0 * 32) / w; };
		printf("%d warp(s)/SMSP, SMSP cycles per element: ffma+ex2+add+cvt %.2f | no add %.2f | no cvt %.2f | all poly %.2f | poly 2/8 %.2f | 3/8 %.2f | 4/8 %.2f\n",
			w, c(4), c(5), c(6), c(7), c(8), c(9), c(10));
		printf("   packed ex2.f16x2 (cycles per ELEMENT): ffma+cvt+ex2h2+hadd2 %.2f | no sum %.2f | f32 sum via 2 cvt %.2f\n",
			hh[4] / (64.0 * 32) / w, hh[5] / (64.0 * 32) / w, hh[6] / (64.0 * 32) / w);
	}
	return 0;
}
