// tools/tmem_rates.cu -- development aid: tensor-memory read / write throughput per SM (tcgen05.ld / tcgen05.st, 32x32b shape,
// the shape the attention softmax warps use). Run on the GPU box.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r)
{
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
		"{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
		: "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
		  "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
		  "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
		  "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
		: "r"(taddr));
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r)
{
	asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
		:: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
		   "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
// MODE 0: loads, MODE 1: stores. `warps` warps per CTA (multiple of 4), one CTA per SM.
template <int MODE> __global__ void k(long long* cyc, float* out, int iters)
{
	__shared__ uint32_t slot;
	const int warp = threadIdx.x >> 5;
	if (warp == 0) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&slot)), "r"(512) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) & 3) * 128;
	uint32_t v[32]; float acc = 0.f;
	for (int i = 0; i < 32; ++i) v[i] = threadIdx.x + i;
	__syncthreads();
	long long t0 = clock64();
	for (int it = 0; it < iters; ++it) {
		if (MODE == 0) {
			tmem_ld32(base, v); tmem_ld32(base + 32, v); tmem_ld32(base + 64, v); tmem_ld32(base + 96, v);
			asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
			acc += __uint_as_float(v[it & 31]);
		} else {
			tmem_st16(base, v); tmem_st16(base + 16, v); tmem_st16(base + 32, v); tmem_st16(base + 48, v);
			tmem_st16(base + 64, v); tmem_st16(base + 80, v); tmem_st16(base + 96, v); tmem_st16(base + 112, v);
			asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
		}
	}
	long long t1 = clock64();
	__syncthreads();
	if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
	out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(slot), "r"(512) : "memory");
}
int main()
{
	long long* cyc; float* out; cudaMalloc(&cyc, 64); cudaMalloc(&out, 1 << 22);
	const int iters = 2000;
	for (int warps : {4, 8, 16}) {
		for (int mode = 0; mode < 2; ++mode) {
			if (mode == 0) k<0><<<148, warps * 32>>>(cyc, out, iters); else k<1><<<148, warps * 32>>>(cyc, out, iters);
			long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
			cudaError_t e = cudaDeviceSynchronize();
			// per iteration every warp moves 4 x (32 lanes x 32 columns x 4 B) = 16 KB
			double bytes = (double)iters * warps * 16384.0;
			printf("%2d warps/CTA %s: %.1f clk per 16 KB warp-iteration, %.1f B/clk per SM (%s)\n", warps, mode ? "tcgen05.st" : "tcgen05.ld",
				(double)h / iters, bytes / (double)h, cudaGetErrorString(e));
		}
	}
	return 0;
}
