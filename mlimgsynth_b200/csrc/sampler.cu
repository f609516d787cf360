// sampler.cu -- device-resident sampler kernels (K7) and image pack/unpack helpers; see
// include/ggml-b200.h for the contract and the reference loops each entry point replaces.
// All HBM-bound, tiny tensors (64-256 KiB per image): one launch per solver stage, float4 accesses.
#include "engine.h"
#include "ggml-b200.h"

using namespace b200;

cudaStream_t b200_engine_stream();   // backend.cpp
namespace b200 { bool g_dryrun(); }
#define DRY (b200::g_dryrun())

namespace {

struct LinArgs {
	float* out[GGML_B200_LINCOMB_MAX_OUT];
	const float* in[GGML_B200_LINCOMB_MAX_IN];
	float coef[GGML_B200_LINCOMB_MAX_OUT * GGML_B200_LINCOMB_MAX_IN];
	int n_out, n_in;
};

template <int VEC>
__global__ void lincomb_kernel(LinArgs a, long long n)
{
	long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * VEC;
	if (i >= n) return;
	float v[GGML_B200_LINCOMB_MAX_IN][VEC];
	#pragma unroll
	for (int k = 0; k < GGML_B200_LINCOMB_MAX_IN; ++k) {
		if (k < a.n_in) {
			if (VEC == 4) { float4 t = *reinterpret_cast<const float4*>(a.in[k] + i); v[k][0] = t.x; v[k][1] = t.y; v[k][2] = t.z; v[k][3] = t.w; }
			else v[k][0] = a.in[k][i];
		}
	}
	#pragma unroll
	for (int j = 0; j < GGML_B200_LINCOMB_MAX_OUT; ++j) {
		if (j < a.n_out) {
			float r[VEC];
			#pragma unroll
			for (int e = 0; e < VEC; ++e) r[e] = 0.f;
			#pragma unroll
			for (int k = 0; k < GGML_B200_LINCOMB_MAX_IN; ++k)
				if (k < a.n_in) {
					float c = a.coef[j * a.n_in + k];
					#pragma unroll
					for (int e = 0; e < VEC; ++e) r[e] += c * v[k][e];   // same left-to-right order as the reference's scalar loops
				}
			if (VEC == 4) *reinterpret_cast<float4*>(a.out[j] + i) = make_float4(r[0], r[1], r[2], r[3]);
			else a.out[j][i] = r[0];
		}
	}
}

__global__ void affine_kernel(float* out, const float* in, float pre, float mul, float post, long long n)
{
	long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
	if (i < n) out[i] = __fadd_rn(__fmul_rn(__fadd_rn(in[i], pre), mul), post);   // no FMA contraction: keep the reference's rounding order
}

__global__ void mask_blend_kernel(float* x, const float* x0, const float* m, long long n_pix, long long total)
{
	long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
	if (i >= total) return;
	float mm = m[i % n_pix];
	x[i] = x0[i] * mm + x[i] * (1 - mm);
}

__global__ void vae_sample_kernel(float* out, const float* mean, const float* logvar, const float* noise, float scale, long long n)
{
	long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
	if (i >= n) return;
	float lv = fminf(fmaxf(logvar[i], -30.f), 20.f);
	out[i] = (mean[i] + (float)exp((double)lv * 0.5) * noise[i]) * scale;
}

__global__ void pack_rgb8_kernel(uint8_t* out, const float* in, int w, int h, int c, float mul, float add)
{
	long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
	long long np = (long long)w * h;
	if (p >= np) return;
	for (int k = 0; k < c; ++k) {
		float v = (in[k * np + p] * mul + add) * 255.f;
		v = fminf(fmaxf(v, 0.f), 255.f);
		out[p * c + k] = (uint8_t)v;   // truncation, as mlimgsynth.c:123-125
	}
}

__global__ void unpack_u8_kernel(float* out, const uint8_t* in, int w, int h, int c_in, int c_first, int c_count, float mul, float add)
{
	long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
	long long np = (long long)w * h;
	if (p >= np) return;
	for (int k = 0; k < c_count; ++k)
		out[k * np + p] = (in[p * c_in + c_first + k] * (1.0f / 255.0f)) * mul + add;
}

__global__ void box_downsample_kernel(float* out, const float* in, int w, int h, int planes, int fw, int fh)
{
	int ow = w / fw, oh = h / fh;
	long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
	long long total = (long long)ow * oh * planes;
	if (i >= total) return;
	int x = (int)(i % ow), y = (int)((i / ow) % oh), pl = (int)(i / ((long long)ow * oh));
	float s = 0.f;
	for (int dy = 0; dy < fh; ++dy)
		for (int dx = 0; dx < fw; ++dx)
			s += in[((long long)pl * h + (y * fh + dy)) * w + x * fw + dx];
	out[i] = s / (fw * fh);
}

// Ordered merge of decoded VAE tiles (vae.c:365-387): every output pixel takes the LAST tile of the reference's row-major list whose kept
// region covers it (later tiles overwrite earlier ones), then out = (v + pre_add) * mul. One launch instead of one 2-D copy per tile and plane.
struct TileMergeArgs {
	int ow, oh, planes, nt0, nt1, tw, th, step0, step1, n0f, n1f, kf, full0, full1, world, slots;
	long long tile_elems; float pre_add, mul;
};
__host__ __device__ inline void tile_merge_pixel(const TileMergeArgs& a, const float* tiles, float* out, int x, int y)
{
	// candidate tile columns / rows: walk backwards, the first hit is the last writer
	int t0 = -1, t1 = -1, sx = 0, sy = 0;
	for (int t = a.nt0 - 1; t >= 0 && t0 < 0; --t) {
		int i0 = t * a.step0; if (i0 > a.ow - a.n0f) i0 = a.ow - a.n0f;
		const int d0 = i0 ? a.kf : 0, c0 = a.full0 ? a.n0f : a.n0f - a.kf;
		if (x >= i0 + d0 && x < i0 + d0 + c0) { t0 = t; sx = x - i0; }
	}
	for (int t = a.nt1 - 1; t >= 0 && t1 < 0; --t) {
		int i1 = t * a.step1; if (i1 > a.oh - a.n1f) i1 = a.oh - a.n1f;
		const int d1 = i1 ? a.kf : 0, c1 = a.full1 ? a.n1f : a.n1f - a.kf;
		if (y >= i1 + d1 && y < i1 + d1 + c1) { t1 = t; sy = y - i1; }
	}
	if (t0 < 0 || t1 < 0) return;          // (not covered: cannot happen for a valid plan; the pixel keeps its value)
	const int t = t1 * a.nt0 + t0;
	const float* src = tiles + a.tile_elems * ((long long)(t % a.world) * a.slots + t / a.world);
	for (int p = 0; p < a.planes; ++p)
		out[((long long)p * a.oh + y) * a.ow + x] = (src[((long long)p * a.th + sy) * a.tw + sx] + a.pre_add) * a.mul;
}
__global__ void tile_merge_kernel(TileMergeArgs a, const float* __restrict__ tiles, float* __restrict__ out)
{
	const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
	if (x < a.ow) tile_merge_pixel(a, tiles, out, x, y);
}

__device__ __forceinline__ float lora_ld(const __half* p) { return __half2float(*p); }
__device__ __forceinline__ float lora_ld(const float* p) { return *p; }
__device__ __forceinline__ void lora_st(__half* p, float v) { *p = __float2half_rn(v); }
__device__ __forceinline__ void lora_st(float* p, float v) { *p = v; }
// W += scale * up . down in the weight type T (lora.c:46-78: operands in wtype, product and sum in f32, ONE rounding back to wtype)
template <typename T>
__global__ void lora_merge_kernel(T* W, const T* down, const T* up, long long n0, long long n1, int r, float scale)
{
	long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;   // input-feature index (contiguous)
	long long j = blockIdx.y;                                         // output-feature index
	if (i >= n0) return;
	float acc = 0.f;
	for (int k = 0; k < r; ++k) acc += lora_ld(up + j * r + k) * lora_ld(down + (long long)k * n0 + i);
	float w = lora_ld(W + j * n0 + i);
	lora_st(W + j * n0 + i, __fadd_rn(w, __fmul_rn(acc, scale)));
}

__global__ void nonfinite_kernel(const float* x, long long n, int* flag)
{
	long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
	bool bad = i < n && !isfinite(x[i]);
	if (__any_sync(~0u, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}

int* g_flag = nullptr;
int* flag_ptr()
{
	if (!g_flag) { CUDA_CHECK(cudaMalloc(&g_flag, sizeof(int))); CUDA_CHECK(cudaMemsetAsync(g_flag, 0, sizeof(int), b200_engine_stream())); }
	return g_flag;
}
inline unsigned nblk(long long n, int t) { return (unsigned)((n + t - 1) / t); }

}  // namespace

extern "C" {

void* ggml_b200_malloc(size_t bytes) { if (DRY) return calloc(1, bytes ? bytes : 4); void* p = nullptr; CUDA_CHECK(cudaMalloc(&p, bytes ? bytes : 4)); return p; }
void ggml_b200_free(void* dev) { if (DRY) { free(dev); return; } if (dev) { cudaStreamSynchronize(b200_engine_stream()); cudaFree(dev); } }
// page-locked host memory for the buffers that cross PCIe every generation (the RGB8 images): D2H at link speed, not at the
// pageable-copy speed
void* ggml_b200_host_malloc(size_t bytes)
{ if (DRY) return malloc(bytes ? bytes : 4); void* p = nullptr; CUDA_CHECK(cudaMallocHost(&p, bytes ? bytes : 4)); return p; }
void ggml_b200_host_free(void* p) { if (!p) return; if (DRY) { free(p); return; } cudaFreeHost(p); }
void ggml_b200_upload(void* dev, const void* host, size_t bytes)
{ if (DRY) { memcpy(dev, host, bytes); return; } CUDA_CHECK(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, b200_engine_stream())); g_stats.h2d_bytes += bytes; }
void ggml_b200_download(void* host, const void* dev, size_t bytes)
{
	if (DRY) { memcpy(host, dev, bytes); return; }
	CUDA_CHECK(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, b200_engine_stream()));
	cudaError_t e = cudaStreamSynchronize(b200_engine_stream());
	if (e != cudaSuccess) B200_FATAL("device work failed: %s", cudaGetErrorString(e));
	g_stats.d2h_bytes += bytes;
}
// Dry-run mode (no GPU: "device" pointers are host allocations) still moves bytes for the pure data-movement helpers, so the
// host-side tile split / merge logic can be exercised by the CPU tests; nothing that computes runs there.
void ggml_b200_copy(void* dst, const void* src, size_t bytes)
{ if (DRY) { memmove(dst, src, bytes); return; } CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, b200_engine_stream())); }
void ggml_b200_memset(void* dev, int value, size_t bytes) { if (DRY) { memset(dev, value, bytes); return; } CUDA_CHECK(cudaMemsetAsync(dev, value, bytes, b200_engine_stream())); }
void ggml_b200_copy2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width_bytes, size_t rows)
{
	if (DRY) { for (size_t r = 0; r < rows; ++r) memmove((char*)dst + r * dpitch, (const char*)src + r * spitch, width_bytes); return; }
	CUDA_CHECK(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width_bytes, rows, cudaMemcpyDeviceToDevice, b200_engine_stream()));
}
void ggml_b200_affine(float* out, const float* in, float pre_add, float mul, float post_add, int64_t n)
{
	if (DRY) return;
	affine_kernel<<<nblk(n, 256), 256, 0, b200_engine_stream()>>>(out, in, pre_add, mul, post_add, n);
	g_stats.kernel_launches++;
}
void ggml_b200_synchronize(void)
{
	if (DRY) return;
	cudaError_t e = cudaStreamSynchronize(b200_engine_stream());
	if (e != cudaSuccess) B200_FATAL("device work failed: %s", cudaGetErrorString(e));
}

void ggml_b200_lincomb(int n_out, float* const* outs, int n_in, const float* const* ins, const float* coef, int64_t n)
{
	if (DRY) return;
	if (n_out < 1 || n_out > GGML_B200_LINCOMB_MAX_OUT || n_in < 1 || n_in > GGML_B200_LINCOMB_MAX_IN) B200_FATAL("ggml_b200_lincomb: bad arity");
	LinArgs a; memset(&a, 0, sizeof(a));
	a.n_out = n_out; a.n_in = n_in;
	bool vec = (n % 4) == 0;
	for (int j = 0; j < n_out; ++j) { a.out[j] = outs[j]; vec = vec && ((uintptr_t)outs[j] % 16 == 0); }
	for (int i = 0; i < n_in; ++i) { a.in[i] = ins[i]; vec = vec && ((uintptr_t)ins[i] % 16 == 0); }
	for (int k = 0; k < n_out * n_in; ++k) a.coef[k] = coef[k];
	if (vec) lincomb_kernel<4><<<nblk(n / 4, 256), 256, 0, b200_engine_stream()>>>(a, n);
	else lincomb_kernel<1><<<nblk(n, 256), 256, 0, b200_engine_stream()>>>(a, n);
	g_stats.kernel_launches++;
}

void ggml_b200_mask_blend(float* x, const float* x0, const float* mask, int64_t n_pix, int64_t n_planes)
{
	if (DRY) return;
	long long total = n_pix * n_planes;
	mask_blend_kernel<<<nblk(total, 256), 256, 0, b200_engine_stream()>>>(x, x0, mask, n_pix, total);
	g_stats.kernel_launches++;
}
void ggml_b200_vae_sample(float* out, const float* mean, const float* logvar, const float* noise, float scale, int64_t n)
{
	if (DRY) return;
	vae_sample_kernel<<<nblk(n, 256), 256, 0, b200_engine_stream()>>>(out, mean, logvar, noise, scale, n);
	g_stats.kernel_launches++;
}
void ggml_b200_pack_rgb8(uint8_t* out_hwc, const float* in_chw, int w, int h, int c, float mul, float add)
{
	if (DRY) return;
	pack_rgb8_kernel<<<nblk((long long)w * h, 256), 256, 0, b200_engine_stream()>>>(out_hwc, in_chw, w, h, c, mul, add);
	g_stats.kernel_launches++;
}
void ggml_b200_unpack_u8(float* out_chw, const uint8_t* in_hwc, int w, int h, int c_in, int c_first, int c_count, float mul, float add)
{
	if (DRY) return;
	unpack_u8_kernel<<<nblk((long long)w * h, 256), 256, 0, b200_engine_stream()>>>(out_chw, in_hwc, w, h, c_in, c_first, c_count, mul, add);
	g_stats.kernel_launches++;
}
void ggml_b200_box_downsample(float* out, const float* in, int w, int h, int planes, int fw, int fh)
{
	if (DRY) return;
	long long total = (long long)(w / fw) * (h / fh) * planes;
	box_downsample_kernel<<<nblk(total, 256), 256, 0, b200_engine_stream()>>>(out, in, w, h, planes, fw, fh);
	g_stats.kernel_launches++;
}
void ggml_b200_tile_merge(float* out, const float* tiles, int ow, int oh, int planes, int nt0, int nt1, int tw, int th, int step0, int step1,
	int keep_margin, int full0, int full1, int world, int slots, float pre_add, float mul)
{
	TileMergeArgs a = { ow, oh, planes, nt0, nt1, tw, th, step0, step1, tw, th, keep_margin, full0, full1, world, slots, (long long)planes * tw * th, pre_add, mul };
	if (DRY) {          // host buffers: the same per-pixel rule (CPU tests of the merge order)
		for (int y = 0; y < oh; ++y) for (int x = 0; x < ow; ++x) tile_merge_pixel(a, tiles, out, x, y);
		return;
	}
	dim3 grid(nblk(ow, 256), (unsigned)oh);
	tile_merge_kernel<<<grid, 256, 0, b200_engine_stream()>>>(a, tiles, out);
	g_stats.kernel_launches++;
}
void ggml_b200_lora_merge_f16(void* w_dev, const void* down_dev, const void* up_dev, int64_t n0, int64_t n1, int r, float scale)
{
	if (DRY) return;
	dim3 grid(nblk(n0, 256), (unsigned)n1);
	lora_merge_kernel<__half><<<grid, 256, 0, b200_engine_stream()>>>((__half*)w_dev, (const __half*)down_dev, (const __half*)up_dev, n0, n1, r, scale);
	g_stats.kernel_launches++;
}
void ggml_b200_lora_merge_f32(void* w_dev, const void* down_dev, const void* up_dev, int64_t n0, int64_t n1, int r, float scale)
{
	if (DRY) return;
	dim3 grid(nblk(n0, 256), (unsigned)n1);
	lora_merge_kernel<float><<<grid, 256, 0, b200_engine_stream()>>>((float*)w_dev, (const float*)down_dev, (const float*)up_dev, n0, n1, r, scale);
	g_stats.kernel_launches++;
}
void ggml_b200_nonfinite_accumulate(const float* x, int64_t n)
{
	if (DRY) return;
	nonfinite_kernel<<<nblk(n, 256), 256, 0, b200_engine_stream()>>>(x, n, flag_ptr());
	g_stats.kernel_launches++;
}
int ggml_b200_nonfinite_check(void)
{
	if (DRY) return 0;
	int v = 0;
	ggml_b200_download(&v, flag_ptr(), sizeof(int));
	if (v) CUDA_CHECK(cudaMemsetAsync(g_flag, 0, sizeof(int), b200_engine_stream()));
	return v;
}
void ggml_b200_profile_enable(int on) { profile_enable(on != 0); }
int ggml_b200_profile_get(int kind, double* ms, double* flops, double* bytes, uint64_t* launches)
{ return profile_get(kind, ms, flops, bytes, launches) ? 1 : 0; }
static cudaEvent_t g_t0 = nullptr, g_t1 = nullptr;
void ggml_b200_timer_start(void)
{
	if (DRY) return;
	if (!g_t0) { CUDA_CHECK(cudaEventCreate(&g_t0)); CUDA_CHECK(cudaEventCreate(&g_t1)); }
	CUDA_CHECK(cudaEventRecord(g_t0, b200_engine_stream()));
}
double ggml_b200_timer_stop(void)
{
	if (DRY || !g_t0) return 0;
	CUDA_CHECK(cudaEventRecord(g_t1, b200_engine_stream()));
	CUDA_CHECK(cudaEventSynchronize(g_t1));
	float ms = 0; CUDA_CHECK(cudaEventElapsedTime(&ms, g_t0, g_t1));
	return ms;
}
const struct ggml_b200_stats* ggml_b200_get_stats(void) { return (const struct ggml_b200_stats*)&g_stats; }
int ggml_b200_last_plan_count(const char* key) { return b200::last_plan_count(key); }

}  // extern "C"
