// backend.cpp -- backend registry, graph allocator, tensor upload/download and graph compute
// entry points of the ggml-shaped C ABI (include/ggml-backend.h, include/ggml-alloc.h).
//
// There is exactly one backend: the B200 CUDA engine. There is no CPU backend and no fallback:
// if no usable device exists, backend init returns NULL and says why.
#include "engine.h"
#include <mutex>

using namespace b200;

namespace b200 { Stats g_stats; }

static ggml_backend_buffer g_dev_buffer = { false };
struct ggml_backend_buffer_type { int device; };
struct ggml_backend_reg { int dummy; };
struct ggml_backend_device { int index; char name[32]; char desc[256]; };

static ggml_backend_reg g_reg;
static std::vector<ggml_backend_device> g_devs;
static std::vector<ggml_backend_buffer_type> g_bufts;
static std::once_flag g_devs_once;

static void enumerate_devices()
{
	std::call_once(g_devs_once, [] {
		int n = 0;
		cudaError_t e = cudaGetDeviceCount(&n);
		if (e != cudaSuccess) { n = 0; cudaGetLastError(); }
		g_devs.resize(n); g_bufts.resize(n);
		for (int i = 0; i < n; ++i) {
			cudaDeviceProp pr;
			g_devs[i].index = i; g_bufts[i].device = i;
			snprintf(g_devs[i].name, sizeof(g_devs[i].name), "CUDA%d", i);
			if (cudaGetDeviceProperties(&pr, i) == cudaSuccess)
				snprintf(g_devs[i].desc, sizeof(g_devs[i].desc), "%s (sm_%d%d, %d SMs)", pr.name, pr.major, pr.minor, pr.multiProcessorCount);
			else snprintf(g_devs[i].desc, sizeof(g_devs[i].desc), "unknown");
		}
	});
}

// GGML_B200_DRYRUN=1: plan graphs without a GPU (host pointers, nothing executes). Used on the
// GPU-less build box to validate the planner; results are undefined, so it is never a compute path.
static bool dryrun() { static int v = -1; if (v < 0) { const char* e = getenv("GGML_B200_DRYRUN"); v = e && *e == '1'; } return v; }
namespace b200 { bool g_dryrun() { return dryrun(); } }

static cudaStream_t g_stream = nullptr;   // tensor_set/get have no backend argument: one engine stream per process
static int g_device = -1;

cudaStream_t b200_engine_stream()
{
	if (!g_stream) B200_FATAL("no B200 backend initialised (call ggml_backend_init_* first)");
	return g_stream;
}

extern "C" {

ggml_backend_t ggml_backend_init_by_name(const char* name, const char* params)
{
	(void)params;
	if (dryrun()) {
		ggml_backend* b = new ggml_backend();
		b->be.device = 0; b->be.sm_count = 148; b->be.name = "B200:dryrun";
		g_devs.resize(1); g_bufts.resize(1);
		return b;
	}
	enumerate_devices();
	int dev = 0;
	if (name && *name) {
		// accepted: "B200", "CUDA", "GPU" (device LOCAL_RANK or 0), "B200:<i>", "CUDA:<i>", "CUDA<i>"
		if (name[0] == 'C' && name[1] == 'P') {
			B200_LOG("backend '%s' requested, but this library has no CPU backend (B200 engine only)", name);
			return nullptr;
		}
		const char* colon = strchr(name, ':');
		bool explicit_dev = false;
		if (colon && colon[1] >= '0' && colon[1] <= '9') { dev = atoi(colon + 1); explicit_dev = true; }
		else if (!strncmp(name, "CUDA", 4) && name[4] >= '0' && name[4] <= '9') { dev = atoi(name + 4); explicit_dev = true; }
		if (!explicit_dev) if (const char* e = getenv("LOCAL_RANK")) dev = atoi(e);
	} else if (const char* e = getenv("LOCAL_RANK")) dev = atoi(e);
	if (g_devs.empty()) { B200_LOG("no CUDA device visible: the B200 engine cannot start (there is no CPU fallback)"); return nullptr; }
	if (dev < 0 || dev >= (int)g_devs.size()) { B200_LOG("device index %d out of range (%zu devices)", dev, g_devs.size()); return nullptr; }
	cudaDeviceProp pr;
	CUDA_CHECK(cudaGetDeviceProperties(&pr, dev));
	if (pr.major != 10) {
		B200_LOG("device %d is sm_%d%d; this engine contains sm_100a code only", dev, pr.major, pr.minor);
		return nullptr;
	}
	CUDA_CHECK(cudaSetDevice(dev));
	// The tcgen05 kernels run with the maximum shared-memory carve-out (up to 227 KB per CTA). Give every other kernel the
	// same preference so that the SMs do not re-partition L1 / shared memory between consecutive kernels of a graph
	// (the streaming element-wise kernels do not depend on L1 capacity). GGML_B200_CARVEOUT=0 keeps the driver default.
	{ const char* e = getenv("GGML_B200_CARVEOUT"); if (!(e && atoi(e) == 0)) CUDA_CHECK(cudaDeviceSetCacheConfig(cudaFuncCachePreferShared)); }
	ggml_backend* b = new ggml_backend();
	b->be.device = dev;
	b->be.sm_count = pr.multiProcessorCount;
	b->be.name = std::string("B200:") + std::to_string(dev);
	if (!g_stream || g_device != dev) {
		CUDA_CHECK(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking));
		g_device = dev;
	}
	b->be.stream = g_stream;
	return b;
}
ggml_backend_t ggml_backend_init_best(void) { return ggml_backend_init_by_name(nullptr, nullptr); }
void ggml_backend_free(ggml_backend_t b) { delete b; }
const char* ggml_backend_name(ggml_backend_t b) { return b->be.name.c_str(); }
ggml_backend_buffer_type_t ggml_backend_get_default_buffer_type(ggml_backend_t b)
{ enumerate_devices(); return &g_bufts[b->be.device]; }
ggml_backend_dev_t ggml_backend_get_device(ggml_backend_t b) { enumerate_devices(); return &g_devs[b->be.device]; }
bool ggml_backend_buffer_is_host(ggml_backend_buffer_t buffer) { return buffer ? buffer->is_host : false; }

size_t ggml_backend_reg_count(void) { return 1; }
ggml_backend_reg_t ggml_backend_reg_get(size_t i) { return i == 0 ? &g_reg : nullptr; }
const char* ggml_backend_reg_name(ggml_backend_reg_t) { return "B200"; }
size_t ggml_backend_reg_dev_count(ggml_backend_reg_t) { enumerate_devices(); return g_devs.size(); }
ggml_backend_dev_t ggml_backend_reg_dev_get(ggml_backend_reg_t, size_t i)
{ enumerate_devices(); return i < g_devs.size() ? &g_devs[i] : nullptr; }
const char* ggml_backend_dev_name(ggml_backend_dev_t d) { return d->name; }
const char* ggml_backend_dev_description(ggml_backend_dev_t d) { return d->desc; }
void ggml_backend_dev_memory(ggml_backend_dev_t d, size_t* free_b, size_t* total_b)
{
	int cur = 0; cudaGetDevice(&cur);
	cudaSetDevice(d->index);
	if (cudaMemGetInfo(free_b, total_b) != cudaSuccess) { *free_b = 0; *total_b = 0; cudaGetLastError(); }
	cudaSetDevice(cur);
}
ggml_backend_reg_t ggml_backend_dev_backend_reg(ggml_backend_dev_t) { return &g_reg; }

// Extra entry points reachable through the registry (the ggml way of exposing backend extras):
//   "ggml_backend_set_n_threads" -> NULL (no host threads to configure; mlimgsynth.c:1120-1128 tolerates NULL)
//   "ggml_b200_stats"            -> const b200::Stats* (*)(void)
static const void* stats_get(void) { return &g_stats; }
void* ggml_backend_reg_get_proc_address(ggml_backend_reg_t, const char* name)
{
	if (!strcmp(name, "ggml_b200_stats")) return (void*)stats_get;
	return nullptr;
}

// ---------------------------------------------------------------- allocator
// Leaves (parameters, inputs) and OUTPUT-flagged nodes get device storage in the logical ggml
// layout, because the caller reads/writes them through ggml_backend_tensor_set/get.
ggml_gallocr_t ggml_gallocr_new(ggml_backend_buffer_type_t) { return new ggml_gallocr(); }

static void galloc_release(ggml_gallocr_t a)
{
	if (a->base) {
		if (dryrun()) free(a->base);
		else {
			if (g_stream) cudaStreamSynchronize(g_stream);
			CUDA_CHECK(cudaFree(a->base));
		}
	}
	a->base = nullptr; a->size = 0;
}
void ggml_gallocr_free(ggml_gallocr_t a)
{
	if (!a) return;
	if (a->graph && a->graph->plan) { plan_free(a->graph->plan); a->graph->plan = nullptr; }
	galloc_release(a);
	delete a;
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

bool ggml_gallocr_reserve(ggml_gallocr_t a, struct ggml_cgraph* g)
{
	if (a->graph && a->graph->plan) { plan_free(a->graph->plan); a->graph->plan = nullptr; }
	galloc_release(a);
	a->graph = g;
	// an OUTPUT alias keeps its storage root addressable
	for (ggml_tensor* t : g->seen)
		if (t->view_src && (t->flags & GGML_TENSOR_FLAG_OUTPUT)) t->view_src->flags |= GGML_TENSOR_FLAG_OUTPUT;
	size_t total = 0;
	std::vector<std::pair<ggml_tensor*, size_t>> offs;
	for (ggml_tensor* t : g->seen) {
		if (t->view_src) continue;
		if (t->op == GGML_OP_NONE || (t->flags & GGML_TENSOR_FLAG_OUTPUT)) {
			offs.push_back({t, total});
			total += align_up(ggml_nbytes(t), 256);
		}
	}
	if (total == 0) total = 256;
	cudaError_t e = cudaSuccess;
	if (dryrun()) a->base = malloc(total); else e = cudaMalloc(&a->base, total);
	if (e != cudaSuccess) {
		B200_LOG("could not allocate %.1f MiB of device memory: %s", total / 1048576.0, cudaGetErrorString(e));
		cudaGetLastError();
		a->base = nullptr;
		return false;
	}
	a->size = total;
	for (auto& pr : offs) { pr.first->data = (char*)a->base + pr.second; pr.first->buffer = &g_dev_buffer; }
	for (ggml_tensor* t : g->seen)
		if (t->view_src && t->view_src->data) { t->data = (char*)t->view_src->data + t->view_offs; t->buffer = &g_dev_buffer; }
	return true;
}
bool ggml_gallocr_alloc_graph(ggml_gallocr_t a, struct ggml_cgraph* g)
{
	if (a->graph != g || !a->base) return ggml_gallocr_reserve(a, g);
	return true;
}
size_t ggml_gallocr_get_buffer_size(ggml_gallocr_t a, int) { return a->size; }

// ---------------------------------------------------------------- upload / download / compute
void ggml_backend_tensor_set(struct ggml_tensor* t, const void* data, size_t offset, size_t size)
{
	if (!t->data) GGML_ABORT("ggml_backend_tensor_set: tensor '%s' has no device storage (not a leaf/output, or graph not allocated)", t->name);
	if (offset + size > ggml_nbytes(t)) GGML_ABORT("ggml_backend_tensor_set: out of bounds write to '%s'", t->name);
	if (dryrun()) { memcpy((char*)t->data + offset, data, size); trec(storage_root(t))->version++; return; }
	// pageable source: the copy is staged before the call returns, so the caller may reuse `data`
	CUDA_CHECK(cudaMemcpyAsync((char*)t->data + offset, data, size, cudaMemcpyHostToDevice, g_stream));
	trec(storage_root(t))->version++;
	g_stats.h2d_bytes += size;
}

void ggml_backend_tensor_get(const struct ggml_tensor* t, void* data, size_t offset, size_t size)
{
	if (!t->data) GGML_ABORT("ggml_backend_tensor_get: tensor '%s' has no device storage (flag it with ggml_set_output)", t->name);
	if (offset + size > ggml_nbytes(t)) GGML_ABORT("ggml_backend_tensor_get: out of bounds read of '%s'", t->name);
	if (dryrun()) { memset(data, 0, size); return; }
	CUDA_CHECK(cudaMemcpyAsync(data, (const char*)t->data + offset, size, cudaMemcpyDeviceToHost, g_stream));
	cudaError_t e = cudaStreamSynchronize(g_stream);
	if (e != cudaSuccess) B200_FATAL("graph execution failed: %s (%s)", cudaGetErrorName(e), cudaGetErrorString(e));
	g_stats.d2h_bytes += size;
}

enum ggml_status ggml_backend_graph_compute(ggml_backend_t backend, struct ggml_cgraph* g)
{
	if (!g->plan || g->built_n_nodes != g->nodes.size()) {
		if (g->plan) plan_free(g->plan);
		g->plan = plan_build(&backend->be, g);
		g->built_n_nodes = g->nodes.size();
		if (!g->plan) return GGML_STATUS_FAILED;
	}
	if (dryrun()) return GGML_STATUS_SUCCESS;
	plan_run(g->plan);
	if (getenv("GGML_B200_SYNC")) {
		cudaError_t e = cudaStreamSynchronize(backend->be.stream);
		if (e != cudaSuccess) { B200_LOG("graph execution failed: %s", cudaGetErrorString(e)); return GGML_STATUS_FAILED; }
	}
	return GGML_STATUS_SUCCESS;
}

}  // extern "C"
