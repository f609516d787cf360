// ggml_api.cpp -- graph recorder: contexts, tensors, the 30 op builders, graph expansion,
// metadata and the host conversion helpers of the ggml-shaped C ABI (include/ggml.h).
// Nothing here computes: builders only record nodes (shape inference + op params). The
// recorded graph is planned and executed by backend.cpp / planner.cpp.
#include "engine.h"
#include <cmath>
#include <cstdarg>
#include <unordered_set>

using namespace b200;

extern "C" {

void ggml_abort(const char* file, int line, const char* fmt, ...)
{
	va_list ap;
	fprintf(stderr, "[ggml_b200] %s:%d: ", file, line);
	va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap);
	fputc('\n', stderr);
	abort();
}

// ---------------------------------------------------------------- types
struct TypeInfo { const char* name; size_t size; };
static TypeInfo type_info(enum ggml_type t)
{
	switch (t) {
	case GGML_TYPE_F32:  return {"f32", 4};
	case GGML_TYPE_F16:  return {"f16", 2};
	case GGML_TYPE_BF16: return {"bf16", 2};
	case GGML_TYPE_I8:   return {"i8", 1};
	case GGML_TYPE_I16:  return {"i16", 2};
	case GGML_TYPE_I32:  return {"i32", 4};
	case GGML_TYPE_I64:  return {"i64", 8};
	case GGML_TYPE_F64:  return {"f64", 8};
	default:             return {nullptr, 0};
	}
}

size_t ggml_type_size(enum ggml_type t)
{
	TypeInfo ti = type_info(t);
	if (!ti.size) GGML_ABORT("tensor type %d is not supported by the B200 engine (f32/f16/bf16/int only)", (int)t);
	return ti.size;
}
const char* ggml_type_name(enum ggml_type t)
{
	TypeInfo ti = type_info(t);
	return ti.name ? ti.name : "unsupported";
}

float ggml_fp16_to_fp32(ggml_fp16_t x) { __half h; memcpy(&h, &x, 2); return __half2float(h); }
ggml_fp16_t ggml_fp32_to_fp16(float x) { __half h = __float2half_rn(x); ggml_fp16_t r; memcpy(&r, &h, 2); return r; }
void ggml_fp16_to_fp32_row(const ggml_fp16_t* x, float* y, int64_t n)
{ for (int64_t i = 0; i < n; ++i) y[i] = ggml_fp16_to_fp32(x[i]); }
void ggml_fp32_to_fp16_row(const float* x, ggml_fp16_t* y, int64_t n)
{ for (int64_t i = 0; i < n; ++i) y[i] = ggml_fp32_to_fp16(x[i]); }
void ggml_bf16_to_fp32_row(const ggml_bf16_t* x, float* y, int64_t n)
{
	for (int64_t i = 0; i < n; ++i) { uint32_t u = (uint32_t)x[i].bits << 16; memcpy(&y[i], &u, 4); }
}
size_t ggml_quantize_chunk(enum ggml_type type, const float*, void*, int64_t, int64_t, int64_t, const float*)
{
	B200_LOG("quantised weight type %d requested: not supported by the B200 engine", (int)type);
	return 0;
}
static void tt_f16(const void* x, float* y, int64_t k) { ggml_fp16_to_fp32_row((const ggml_fp16_t*)x, y, k); }
static void tt_bf16(const void* x, float* y, int64_t k) { ggml_bf16_to_fp32_row((const ggml_bf16_t*)x, y, k); }
const struct ggml_type_traits* ggml_get_type_traits(enum ggml_type t)
{
	static ggml_type_traits tr[GGML_TYPE_COUNT];
	GGML_ASSERT((unsigned)t < GGML_TYPE_COUNT);
	TypeInfo ti = type_info(t);
	tr[t].type_name = ggml_type_name(t);
	tr[t].blck_size = 1;
	tr[t].type_size = ti.size;
	tr[t].is_quantized = false;
	tr[t].to_float = t == GGML_TYPE_F16 ? tt_f16 : t == GGML_TYPE_BF16 ? tt_bf16 : nullptr;
	return &tr[t];
}

// ---------------------------------------------------------------- contexts & tensors
struct TensorBlock { ggml_tensor t; TRec rec; };

struct ggml_context* ggml_init(struct ggml_init_params) { return new ggml_context(); }

void ggml_free(struct ggml_context* ctx)
{
	if (!ctx) return;
	for (ggml_cgraph* g : ctx->graphs) {
		if (g->plan) plan_free(g->plan);
		delete g;
	}
	for (ggml_tensor* t : ctx->tensors) delete (TensorBlock*)t;
	delete ctx;
}

size_t ggml_tensor_overhead(void) { return sizeof(TensorBlock); }
size_t ggml_graph_overhead(void) { return sizeof(ggml_cgraph) + 4096; }

static ggml_tensor* new_tensor(ggml_context* ctx, enum ggml_type type, const int64_t ne[4])
{
	TensorBlock* b = new TensorBlock();
	memset(&b->t, 0, sizeof(b->t));
	ggml_tensor* t = &b->t;
	t->type = type;
	t->nb[0] = ggml_type_size(type);
	for (int i = 0; i < 4; ++i) {
		GGML_ASSERT(ne[i] >= 1);
		t->ne[i] = ne[i];
		if (i) t->nb[i] = t->nb[i-1] * t->ne[i-1];
	}
	t->extra = &b->rec;
	ctx->tensors.push_back(t);
	return t;
}

struct ggml_tensor* ggml_new_tensor_1d(struct ggml_context* c, enum ggml_type t, int64_t n0)
{ int64_t ne[4] = {n0, 1, 1, 1}; return new_tensor(c, t, ne); }
struct ggml_tensor* ggml_new_tensor_2d(struct ggml_context* c, enum ggml_type t, int64_t n0, int64_t n1)
{ int64_t ne[4] = {n0, n1, 1, 1}; return new_tensor(c, t, ne); }
struct ggml_tensor* ggml_new_tensor_4d(struct ggml_context* c, enum ggml_type t, int64_t n0, int64_t n1, int64_t n2, int64_t n3)
{ int64_t ne[4] = {n0, n1, n2, n3}; return new_tensor(c, t, ne); }

struct ggml_tensor* ggml_get_first_tensor(const struct ggml_context* ctx)
{ return ctx->tensors.empty() ? nullptr : ctx->tensors[0]; }
struct ggml_tensor* ggml_get_next_tensor(const struct ggml_context* ctx, struct ggml_tensor* t)
{
	for (size_t i = 0; i + 1 < ctx->tensors.size(); ++i) if (ctx->tensors[i] == t) return ctx->tensors[i+1];
	return nullptr;
}

struct ggml_tensor* ggml_set_name(struct ggml_tensor* t, const char* name)
{ snprintf(t->name, sizeof(t->name), "%s", name); return t; }
const char* ggml_get_name(const struct ggml_tensor* t) { return t->name; }
void ggml_set_input(struct ggml_tensor* t) { t->flags |= GGML_TENSOR_FLAG_INPUT; }
void ggml_set_output(struct ggml_tensor* t) { t->flags |= GGML_TENSOR_FLAG_OUTPUT; }
int64_t ggml_nelements(const struct ggml_tensor* t) { return t->ne[0] * t->ne[1] * t->ne[2] * t->ne[3]; }
size_t ggml_element_size(const struct ggml_tensor* t) { return ggml_type_size(t->type); }
size_t ggml_nbytes(const struct ggml_tensor* t)
{
	size_t n = ggml_type_size(t->type);
	for (int i = 0; i < 4; ++i) n += (size_t)(t->ne[i] - 1) * t->nb[i];
	return n;
}
int ggml_n_dims(const struct ggml_tensor* t)
{
	for (int i = 3; i >= 1; --i) if (t->ne[i] > 1) return i + 1;
	return 1;
}

const char* ggml_op_name(enum ggml_op op)
{
	static const char* names[GGML_OP_COUNT] = {
		"NONE", "ADD", "MUL", "SCALE", "NORM", "GROUP_NORM", "MUL_MAT", "CONT", "RESHAPE", "VIEW",
		"PERMUTE", "TRANSPOSE", "GET_ROWS", "DIAG_MASK_INF", "SOFT_MAX", "CONV_2D", "CONCAT", "PAD",
		"UPSCALE", "TIMESTEP_EMBEDDING", "UNARY", "MAP_CUSTOM1" };
	return (unsigned)op < GGML_OP_COUNT ? names[op] : "?";
}
const char* ggml_op_desc(const struct ggml_tensor* t)
{
	static const char* un[GGML_UNARY_OP_COUNT] = { "TANH", "RELU", "GELU", "GELU_QUICK", "SILU" };
	if (t->op == GGML_OP_UNARY) return un[t->op_params[0]];
	return ggml_op_name(t->op);
}

// ---------------------------------------------------------------- op builders
static bool contiguous(const ggml_tensor* t)
{
	size_t nb = ggml_type_size(t->type);
	for (int i = 0; i < 4; ++i) {
		if (t->ne[i] != 1 && t->nb[i] != nb) return false;
		nb *= t->ne[i];
	}
	return true;
}

static ggml_tensor* node(ggml_context* ctx, enum ggml_op op, enum ggml_type type, const int64_t ne[4],
	ggml_tensor* a, ggml_tensor* b = nullptr)
{
	ggml_tensor* t = new_tensor(ctx, type, ne);
	t->op = op; t->src[0] = a; t->src[1] = b;
	return t;
}

// Node whose result aliases a's storage in ggml semantics (views and *_inplace ops).
static ggml_tensor* alias(ggml_context* ctx, enum ggml_op op, ggml_tensor* a, ggml_tensor* b = nullptr)
{
	ggml_tensor* t = node(ctx, op, a->type, a->ne, a, b);
	memcpy(t->nb, a->nb, sizeof(t->nb));
	t->view_src = storage_root(a);
	t->view_offs = a->view_src ? a->view_offs : 0;
	return t;
}

static void check_bcast(const ggml_tensor* a, const ggml_tensor* b)
{
	for (int i = 0; i < 4; ++i)
		if (a->ne[i] % b->ne[i]) GGML_ABORT("operand 2 (%s) cannot be broadcast onto operand 1 (%s)", b->name, a->name);
}

struct ggml_tensor* ggml_add(struct ggml_context* c, struct ggml_tensor* a, struct ggml_tensor* b)
{ check_bcast(a, b); return node(c, GGML_OP_ADD, a->type, a->ne, a, b); }
struct ggml_tensor* ggml_add_inplace(struct ggml_context* c, struct ggml_tensor* a, struct ggml_tensor* b)
{ check_bcast(a, b); return alias(c, GGML_OP_ADD, a, b); }
struct ggml_tensor* ggml_mul(struct ggml_context* c, struct ggml_tensor* a, struct ggml_tensor* b)
{ check_bcast(a, b); return node(c, GGML_OP_MUL, a->type, a->ne, a, b); }

static ggml_tensor* scale_node(ggml_context* c, ggml_tensor* a, float s, bool inplace)
{
	ggml_tensor* t = inplace ? alias(c, GGML_OP_SCALE, a) : node(c, GGML_OP_SCALE, a->type, a->ne, a);
	memcpy(t->op_params, &s, sizeof(s));
	return t;
}
struct ggml_tensor* ggml_scale(struct ggml_context* c, struct ggml_tensor* a, float s) { return scale_node(c, a, s, false); }
struct ggml_tensor* ggml_scale_inplace(struct ggml_context* c, struct ggml_tensor* a, float s) { return scale_node(c, a, s, true); }

struct ggml_tensor* ggml_mul_mat(struct ggml_context* c, struct ggml_tensor* a, struct ggml_tensor* b)
{
	GGML_ASSERT(a->ne[0] == b->ne[0]);
	GGML_ASSERT(b->ne[2] % a->ne[2] == 0 && b->ne[3] % a->ne[3] == 0);
	GGML_ASSERT(a->nb[0] == ggml_type_size(a->type));
	const int64_t ne[4] = { a->ne[1], b->ne[1], b->ne[2], b->ne[3] };
	return node(c, GGML_OP_MUL_MAT, GGML_TYPE_F32, ne, a, b);
}

struct ggml_tensor* ggml_conv_2d(struct ggml_context* c, struct ggml_tensor* a, struct ggml_tensor* b,
	int s0, int s1, int p0, int p1, int d0, int d1)
{
	GGML_ASSERT(a->ne[2] == b->ne[2]);
	const int64_t ne[4] = {
		(b->ne[0] + 2 * p0 - d0 * (a->ne[0] - 1) - 1) / s0 + 1,
		(b->ne[1] + 2 * p1 - d1 * (a->ne[1] - 1) - 1) / s1 + 1,
		a->ne[3], b->ne[3] };
	ggml_tensor* t = node(c, GGML_OP_CONV_2D, GGML_TYPE_F32, ne, a, b);
	const int32_t p[6] = { s0, s1, p0, p1, d0, d1 };
	memcpy(t->op_params, p, sizeof(p));
	return t;
}

struct ggml_tensor* ggml_norm(struct ggml_context* c, struct ggml_tensor* a, float eps)
{
	ggml_tensor* t = node(c, GGML_OP_NORM, a->type, a->ne, a);
	memcpy(t->op_params, &eps, sizeof(eps));
	return t;
}
struct ggml_tensor* ggml_group_norm(struct ggml_context* c, struct ggml_tensor* a, int n_groups, float eps)
{
	ggml_tensor* t = node(c, GGML_OP_GROUP_NORM, a->type, a->ne, a);
	t->op_params[0] = n_groups;
	memcpy(&t->op_params[1], &eps, sizeof(eps));
	return t;
}

static ggml_tensor* unary_node(ggml_context* c, ggml_tensor* a, int op, bool inplace)
{
	ggml_tensor* t = inplace ? alias(c, GGML_OP_UNARY, a) : node(c, GGML_OP_UNARY, a->type, a->ne, a);
	t->op_params[0] = op;
	return t;
}
struct ggml_tensor* ggml_silu(struct ggml_context* c, struct ggml_tensor* a) { return unary_node(c, a, GGML_UNARY_OP_SILU, false); }
struct ggml_tensor* ggml_silu_inplace(struct ggml_context* c, struct ggml_tensor* a) { return unary_node(c, a, GGML_UNARY_OP_SILU, true); }
struct ggml_tensor* ggml_gelu_inplace(struct ggml_context* c, struct ggml_tensor* a) { return unary_node(c, a, GGML_UNARY_OP_GELU, true); }
struct ggml_tensor* ggml_gelu_quick_inplace(struct ggml_context* c, struct ggml_tensor* a) { return unary_node(c, a, GGML_UNARY_OP_GELU_QUICK, true); }
struct ggml_tensor* ggml_relu_inplace(struct ggml_context* c, struct ggml_tensor* a) { return unary_node(c, a, GGML_UNARY_OP_RELU, true); }
struct ggml_tensor* ggml_tanh_inplace(struct ggml_context* c, struct ggml_tensor* a) { return unary_node(c, a, GGML_UNARY_OP_TANH, true); }

struct ggml_tensor* ggml_soft_max_inplace(struct ggml_context* c, struct ggml_tensor* a)
{ return alias(c, GGML_OP_SOFT_MAX, a); }
struct ggml_tensor* ggml_diag_mask_inf_inplace(struct ggml_context* c, struct ggml_tensor* a, int n_past)
{
	ggml_tensor* t = alias(c, GGML_OP_DIAG_MASK_INF, a);
	t->op_params[0] = n_past;
	return t;
}

struct ggml_tensor* ggml_cont(struct ggml_context* c, struct ggml_tensor* a)
{ return node(c, GGML_OP_CONT, a->type, a->ne, a); }

struct ggml_tensor* ggml_permute(struct ggml_context* c, struct ggml_tensor* a, int x0, int x1, int x2, int x3)
{
	const int ax[4] = { x0, x1, x2, x3 };
	int seen = 0;
	for (int i = 0; i < 4; ++i) { GGML_ASSERT(ax[i] >= 0 && ax[i] < 4); seen |= 1 << ax[i]; }
	GGML_ASSERT(seen == 15);
	ggml_tensor* t = alias(c, GGML_OP_PERMUTE, a);
	for (int i = 0; i < 4; ++i) { t->ne[ax[i]] = a->ne[i]; t->nb[ax[i]] = a->nb[i]; }
	memcpy(t->op_params, ax, sizeof(ax));
	return t;
}
struct ggml_tensor* ggml_transpose(struct ggml_context* c, struct ggml_tensor* a)
{
	ggml_tensor* t = alias(c, GGML_OP_TRANSPOSE, a);
	t->ne[0] = a->ne[1]; t->ne[1] = a->ne[0];
	t->nb[0] = a->nb[1]; t->nb[1] = a->nb[0];
	return t;
}

static ggml_tensor* reshape_node(ggml_context* c, ggml_tensor* a, int64_t n0, int64_t n1, int64_t n2, int64_t n3)
{
	GGML_ASSERT(contiguous(a));
	GGML_ASSERT(ggml_nelements(a) == n0 * n1 * n2 * n3);
	ggml_tensor* t = alias(c, GGML_OP_RESHAPE, a);
	const int64_t ne[4] = { n0, n1, n2, n3 };
	t->nb[0] = ggml_type_size(a->type);
	for (int i = 0; i < 4; ++i) { t->ne[i] = ne[i]; if (i) t->nb[i] = t->nb[i-1] * t->ne[i-1]; }
	return t;
}
struct ggml_tensor* ggml_reshape_3d(struct ggml_context* c, struct ggml_tensor* a, int64_t n0, int64_t n1, int64_t n2)
{ return reshape_node(c, a, n0, n1, n2, 1); }
struct ggml_tensor* ggml_reshape_4d(struct ggml_context* c, struct ggml_tensor* a, int64_t n0, int64_t n1, int64_t n2, int64_t n3)
{ return reshape_node(c, a, n0, n1, n2, n3); }

struct ggml_tensor* ggml_view_4d(struct ggml_context* c, struct ggml_tensor* a,
	int64_t n0, int64_t n1, int64_t n2, int64_t n3, size_t nb1, size_t nb2, size_t nb3, size_t offset)
{
	ggml_tensor* t = alias(c, GGML_OP_VIEW, a);
	t->ne[0] = n0; t->ne[1] = n1; t->ne[2] = n2; t->ne[3] = n3;
	t->nb[0] = ggml_type_size(a->type); t->nb[1] = nb1; t->nb[2] = nb2; t->nb[3] = nb3;
	t->view_offs += offset;
	memcpy(t->op_params, &offset, sizeof(offset));
	return t;
}
struct ggml_tensor* ggml_view_1d(struct ggml_context* c, struct ggml_tensor* a, int64_t n0, size_t offset)
{
	const size_t row = ggml_type_size(a->type) * n0;
	return ggml_view_4d(c, a, n0, 1, 1, 1, row, row, row, offset);
}

struct ggml_tensor* ggml_concat(struct ggml_context* c, struct ggml_tensor* a, struct ggml_tensor* b, int dim)
{
	GGML_ASSERT(dim >= 0 && dim < 4);
	int64_t ne[4];
	for (int i = 0; i < 4; ++i) {
		if (i == dim) ne[i] = a->ne[i] + b->ne[i];
		else { GGML_ASSERT(a->ne[i] == b->ne[i]); ne[i] = a->ne[i]; }
	}
	ggml_tensor* t = node(c, GGML_OP_CONCAT, a->type, ne, a, b);
	t->op_params[0] = dim;
	return t;
}
struct ggml_tensor* ggml_pad(struct ggml_context* c, struct ggml_tensor* a, int p0, int p1, int p2, int p3)
{
	const int64_t ne[4] = { a->ne[0] + p0, a->ne[1] + p1, a->ne[2] + p2, a->ne[3] + p3 };
	return node(c, GGML_OP_PAD, a->type, ne, a);
}
struct ggml_tensor* ggml_upscale(struct ggml_context* c, struct ggml_tensor* a, int sf, enum ggml_scale_mode mode)
{
	if (mode != GGML_SCALE_MODE_NEAREST) GGML_ABORT("ggml_upscale: only NEAREST is implemented");
	const int64_t ne[4] = { a->ne[0] * sf, a->ne[1] * sf, a->ne[2], a->ne[3] };
	ggml_tensor* t = node(c, GGML_OP_UPSCALE, a->type, ne, a);
	t->op_params[0] = mode;
	return t;
}
struct ggml_tensor* ggml_get_rows(struct ggml_context* c, struct ggml_tensor* a, struct ggml_tensor* b)
{
	GGML_ASSERT(a->ne[2] == b->ne[1] && b->ne[3] == 1 && b->type == GGML_TYPE_I32);
	const int64_t ne[4] = { a->ne[0], b->ne[0], b->ne[1], b->ne[2] };
	return node(c, GGML_OP_GET_ROWS, GGML_TYPE_F32, ne, a, b);
}
struct ggml_tensor* ggml_timestep_embedding(struct ggml_context* c, struct ggml_tensor* ts, int dim, int max_period)
{
	const int64_t ne[4] = { dim + (dim & 1), ts->ne[0], 1, 1 };
	ggml_tensor* t = node(c, GGML_OP_TIMESTEP_EMBEDDING, GGML_TYPE_F32, ne, ts);
	t->op_params[0] = dim; t->op_params[1] = max_period;
	return t;
}
struct ggml_tensor* ggml_map_custom1_inplace(struct ggml_context*, struct ggml_tensor*, ggml_custom1_op_t, int, void*)
{
	// The reference only reaches this through its CPU-only debug taps (ggml_extend.c:171-198),
	// which are guarded by ggml_backend_buffer_is_host(); device buffers are never host.
	GGML_ABORT("ggml_map_custom1_inplace: host callbacks cannot run on B200 device tensors");
	return nullptr;
}

// ---------------------------------------------------------------- graph
struct ggml_cgraph* ggml_new_graph_custom(struct ggml_context* ctx, size_t size, bool)
{
	static int next_id = 1;
	ggml_cgraph* g = new ggml_cgraph();
	g->size = (int)size;
	g->id = next_id++;
	ctx->graphs.push_back(g);
	return g;
}

static void expand(ggml_cgraph* g, ggml_tensor* t)
{
	// iterative post-order DFS (UNet chains are thousands of nodes deep)
	struct Frame { ggml_tensor* t; int next; };
	std::vector<Frame> stack;
	auto push = [&](ggml_tensor* x) {
		if (trec(x)->seen_graph == g->id) return;
		trec(x)->seen_graph = g->id;
		g->seen.push_back(x);
		stack.push_back({x, 0});
	};
	push(t);
	while (!stack.empty()) {
		Frame& f = stack.back();
		if (f.next < GGML_MAX_SRC) {
			ggml_tensor* s = f.t->src[f.next++];
			if (s) push(s);
		} else {
			ggml_tensor* x = f.t;
			stack.pop_back();
			if (x->op == GGML_OP_NONE) g->leafs.push_back(x);
			else {
				if ((int)g->nodes.size() >= g->size) GGML_ABORT("graph size %d exceeded", g->size);
				g->nodes.push_back(x);
			}
		}
	}
}
void ggml_build_forward_expand(struct ggml_cgraph* g, struct ggml_tensor* t) { expand(g, t); }
int ggml_graph_size(struct ggml_cgraph* g) { return g->size; }
int ggml_graph_n_nodes(struct ggml_cgraph* g) { return (int)g->nodes.size(); }

}  // extern "C"
