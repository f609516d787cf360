// engine.h -- internal declarations shared by the recorder, planner and kernels of the
// B200 engine behind the ggml-shaped C ABI (include/ggml*.h).
#pragma once
#include "ggml.h"
#include "ggml-alloc.h"
#include "ggml-backend.h"

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#define B200_LOG(...) do { fprintf(stderr, "[ggml_b200] " __VA_ARGS__); fputc('\n', stderr); } while (0)
#define B200_FATAL(...) do { B200_LOG(__VA_ARGS__); abort(); } while (0)
#define CUDA_CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) \
	B200_FATAL("CUDA error %s at %s:%d: %s", cudaGetErrorName(e_), __FILE__, __LINE__, cudaGetErrorString(e_)); } while (0)

namespace b200 {

enum DT : int { DT_F32 = 0, DT_F16 = 1, DT_I32 = 2 };
static inline size_t dt_size(DT d) { return d == DT_F16 ? 2 : 4; }

// Per-tensor engine record, allocated together with the ggml_tensor (tensor->extra).
struct TRec {
	uint64_t version = 0;   // bumped by every ggml_backend_tensor_set on this storage root
	int      seen_graph = 0;
};

// A strided device view: element strides, logical ggml dim order (dim 0 first).
struct View {
	void*   ptr = nullptr;
	DT      dt = DT_F32;
	int64_t ne[4] = {1, 1, 1, 1};
	int64_t st[4] = {0, 0, 0, 0};
	int64_t numel() const { return ne[0] * ne[1] * ne[2] * ne[3]; }
};

enum UnaryOp : int { U_TANH = 0, U_RELU, U_GELU, U_GELU_QUICK, U_SILU, U_NONE, U_SCALE };
enum BinOp : int { B_ADD = 0, B_MUL };

struct Plan;
struct Backend;

Plan* plan_build(Backend* be, ggml_cgraph* g);
int last_plan_count(const char* key);      // step histogram of the most recently built plan
void  plan_run(Plan* p);
void  plan_free(Plan* p);
void  profile_enable(bool on);
bool  profile_get(int kind, double* ms, double* flops, double* bytes, uint64_t* launches);

// Counters exposed through ggml_backend_reg_get_proc_address("ggml_b200_stats").
struct Stats {
	uint64_t kernel_launches = 0;   // kernels launched by this library (graph replays counted per node)
	uint64_t graph_launches = 0;
	uint64_t plans_built = 0;
	uint64_t h2d_bytes = 0, d2h_bytes = 0;
};
extern Stats g_stats;

struct Backend {
	int device = 0;
	cudaStream_t stream = nullptr;
	int sm_count = 0;
	std::string name;
};

}  // namespace b200

// ---- C structs behind the opaque handles of the ABI ----
struct ggml_backend_buffer { bool is_host; };

struct ggml_context {
	std::vector<ggml_tensor*> tensors;
	std::vector<ggml_cgraph*> graphs;
};

struct ggml_cgraph {
	int size = 0;
	std::vector<ggml_tensor*> nodes;   // ops, topological order
	std::vector<ggml_tensor*> leafs;   // parameters / inputs
	std::vector<ggml_tensor*> seen;
	b200::Plan* plan = nullptr;
	uint64_t built_n_nodes = 0;
	int id = 0;
};

struct ggml_backend { b200::Backend be; };

struct ggml_gallocr {
	void*  base = nullptr;
	size_t size = 0;
	ggml_cgraph* graph = nullptr;
};

static inline b200::TRec* trec(const ggml_tensor* t) { return (b200::TRec*)t->extra; }
static inline ggml_tensor* storage_root(ggml_tensor* t) { return t->view_src ? t->view_src : t; }
