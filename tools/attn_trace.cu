// tools/attn_trace.cu -- development aid: times attn_tc_kernel on one shape and prints the per-block timeline of
// CTA 0 (softmax warps of both tiles and the MMA thread). Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O2 -I include -I mlimgsynth_b200/csrc tools/attn_trace.cu \
//        -L mlimgsynth_b200/lib -lggml_b200 -Xlinker -rpath=$PWD/mlimgsynth_b200/lib -o mlimgsynth_b200/build/attn_trace
//   mlimgsynth_b200/build/attn_trace 40 4096 4096 8 16
#include "kernels.h"
#include <vector>
#include <random>
using namespace b200;

int main(int argc, char** argv)
{
	int d = argc > 1 ? atoi(argv[1]) : 40, nq = argc > 2 ? atoi(argv[2]) : 4096, nk = argc > 3 ? atoi(argv[3]) : 4096;
	int H = argc > 4 ? atoi(argv[4]) : 8, B = argc > 5 ? atoi(argv[5]) : 16;
	int nshow = argc > 6 ? atoi(argv[6]) : 6;
	size_t nQ = (size_t)B * nq * H * d, nK = (size_t)B * nk * H * d;
	std::vector<__half> hq(nQ), hk(nK), hv(nK);
	std::mt19937 rng(1); std::normal_distribution<float> nd(0.f, 1.f);
	for (auto& x : hq) x = __float2half(nd(rng));
	for (auto& x : hk) x = __float2half(nd(rng));
	for (auto& x : hv) x = __float2half(nd(rng));
	__half *q, *k, *v, *o; long long* tr;
	CUDA_CHECK(cudaMalloc(&q, nQ * 2)); CUDA_CHECK(cudaMalloc(&k, nK * 2)); CUDA_CHECK(cudaMalloc(&v, nK * 2)); CUDA_CHECK(cudaMalloc(&o, nQ * 2));
	CUDA_CHECK(cudaMalloc(&tr, 4 * 64 * 8 * 8)); CUDA_CHECK(cudaMemset(tr, 0, 4 * 64 * 8 * 8));
	CUDA_CHECK(cudaMemcpy(q, hq.data(), nQ * 2, cudaMemcpyHostToDevice));
	CUDA_CHECK(cudaMemcpy(k, hk.data(), nK * 2, cudaMemcpyHostToDevice));
	CUDA_CHECK(cudaMemcpy(v, hv.data(), nK * 2, cudaMemcpyHostToDevice));
	auto mk = [&](void* p, int n, bool vt) {      // token-major [B][n][H][d]
		View w; w.ptr = p; w.dt = DT_F16;
		if (!vt) { w.ne[0] = d; w.ne[1] = n; w.ne[2] = H; w.ne[3] = B; w.st[0] = 1; w.st[1] = (int64_t)H * d; w.st[2] = d; w.st[3] = (int64_t)n * H * d; }
		else     { w.ne[0] = n; w.ne[1] = d; w.ne[2] = H; w.ne[3] = B; w.st[0] = (int64_t)H * d; w.st[1] = 1; w.st[2] = d; w.st[3] = (int64_t)n * H * d; }
		return w;
	};
	View vq = mk(q, nq, false), vk = mk(k, nk, false), vv = mk(v, nk, true), vo = mk(o, nq, false);
	if (!attn_tc_supported(vo, vq, vk, vv, false)) { printf("unsupported\n"); return 1; }
	AttnTC* a = attn_tc_prepare(vo, vq, vk, vv, 1.0f / sqrtf((float)d));
	cudaStream_t s; CUDA_CHECK(cudaStreamCreate(&s));
	for (int i = 0; i < 3; ++i) attn_tc_launch(s, a);
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	cudaEventRecord(e0, s);
	const int n = 10;
	for (int i = 0; i < n; ++i) attn_tc_launch(s, a);
	cudaEventRecord(e1, s); CUDA_CHECK(cudaStreamSynchronize(s));
	float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= n;
	double fl = 4.0 * nq * nk * d * H * B;
	printf("d=%d nq=%d nk=%d H=%d B=%d : %.1f us  %.1f TFLOP/s  %.2f Texp/s\n", d, nq, nk, H, B, ms * 1e3, fl / ms / 1e9, (double)nq * nk * H * B / ms / 1e9);
	attn_tc_set_trace(a, tr);
	attn_tc_launch(s, a); CUDA_CHECK(cudaStreamSynchronize(s));
	std::vector<long long> ht(4 * 64 * 8);
	CUDA_CHECK(cudaMemcpy(ht.data(), tr, ht.size() * 8, cudaMemcpyDeviceToHost));
	if (ht[2047] > 0) printf("clock probe (mid-run CTA): %lld cycles in %lld ns -> SM clock %.0f MHz under this kernel\n", ht[2046], ht[2047], 1e3 * (double)ht[2046] / (double)ht[2047]);
	long long t0 = ht[(2 * 64 + 0) * 8 + 3];     // first QK issue
	auto T = [&](int role, int j, int ev) { long long x = ht[(role * 64 + j) * 8 + ev]; return x ? (long long)(x - t0) : -1; };
	if (getenv("GGML_B200_ATTN_SPLIT") && atoi(getenv("GGML_B200_ATTN_SPLIT")) == 5) {
		// attn_ap_kernel: stream events 0 step start, 1 scores of the next block seen + load issued, 2 first half of the exponentials
		// issued, 3 next block's scores in registers, 4 second half issued, 5 stores complete; issuing warp of the stream:
		// 0 p_full seen, 1 PV issued, 2 QK(u+3) issued
		long long tz = ht[(0 * 64 + 0) * 8 + 0];
		auto R = [&](int role, int j, int ev) { long long x = ht[(role * 64 + j) * 8 + ev]; return x ? (long long)(x - tz) : -1; };
		for (int j = 0; j < nshow; ++j)
			for (int t = 0; t < 2; ++t) {
				printf("blk %2d stream %d:", j, t);
				for (int e = 0; e < 6; ++e) printf(" %7lld", R(t, j, e));
				printf("   mma:");
				for (int e = 0; e < 3; ++e) printf(" %7lld", R(2, j, t * 4 + e));
				printf("\n");
			}
	} else {
	printf("softmax events: 0 wait_s  1 s_ready  2 max_done  3 turn  4 exp_done  5 p_arrived ; mma: 0 pv_wait 1 p_ready 2 pv_issued 3 qk_issued (per tile)\n");
	for (int j = 0; j < nshow; ++j) {
		for (int t = 0; t < 2; ++t) {
			printf("blk %2d tile %d sm:", j, t);
			for (int e = 0; e < 6; ++e) printf(" %7lld", T(t, j, e));
			printf("   mma:");
			for (int e = 0; e < 4; ++e) printf(" %7lld", T(2, j, t * 4 + e));
			printf("\n");
		}
	}
	printf("mma warp detail: 0 before k_full wait  1 after  2 elected (QK issue starts)  3 before v_full wait  4 after  5 p_full[a] seen  6 p_full[b] seen\n");
	for (int j = 1; j < nshow; ++j) { printf("blk %2d mma detail:", j); for (int e = 0; e < 8; ++e) printf(" %7lld", T(3, j, e)); printf("\n"); }
	}
	// check a few outputs against a straightforward CPU evaluation (row 0 and row nq-1 of head 0, image 0)
	std::vector<__half> ho(nQ);
	CUDA_CHECK(cudaMemcpy(ho.data(), o, nQ * 2, cudaMemcpyDeviceToHost));
	double maxerr = 0;
	for (int row : {0, nq / 2 + 3, nq - 1}) {
		std::vector<double> sc(nk); double mx = -1e30;
		for (int j = 0; j < nk; ++j) { double acc = 0; for (int c = 0; c < d; ++c) acc += (double)__half2float(hq[((size_t)row * H) * d + c]) * __half2float(hk[((size_t)j * H) * d + c]); sc[j] = acc / sqrt((double)d); mx = std::max(mx, sc[j]); }
		double sum = 0; for (int j = 0; j < nk; ++j) { sc[j] = exp(sc[j] - mx); sum += sc[j]; }
		for (int c = 0; c < d; ++c) { double acc = 0; for (int j = 0; j < nk; ++j) acc += sc[j] * __half2float(hv[((size_t)j * H) * d + c]); acc /= sum;
			maxerr = std::max(maxerr, fabs(acc - (double)__half2float(ho[((size_t)row * H) * d + c]))); }
	}
	printf("max abs err vs f64 reference on 3 rows: %.3e\n", maxerr);
	return 0;
}
