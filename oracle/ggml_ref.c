/* oracle/ggml_ref.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C CPU restatement of the ggml operator semantics that the reference
 * (aagdev/mlimgsynth) relies on, exported behind the same ggml-shaped C ABI
 * as the product engine (include/ggml.h, ggml-alloc.h, ggml-backend.h). Linked
 * together with the reference's own unmodified host sources (oracle/Makefile)
 * it yields a complete CPU implementation of the reference path that the
 * tests use as the checker and bench.py uses as the CPU baseline.
 *
 * PARITY UNPINNED for the arithmetic: the reference delegates all tensor math
 * to ggml, which is an external, un-vendored, un-pinned dependency that is
 * absent from /root/reference and from this machine (reference Makefile:18-31,
 * README.md:19-25). The reference's own tests pin no numeric output of a ggml
 * graph. The semantics below restate ggml's published CPU
 * behaviour as catalogued in SURVEY.md Appendix A, anchored on the reference's
 * call sites (cited per function). The reference's host logic (tokenizer,
 * prompt parser, Philox RNG, sampler, solvers) is NOT restated here: the
 * oracle uses the reference's own compiled objects for those, and those ARE
 * pinned by the reference's known-answer tests (tests/test_host_cpu.py).
 * What stands in for the missing ggml vectors is a set of independent anchors on the
 * canonical PyTorch definitions of the same operators and models (all CPU tests):
 * tests/test_oracle_vs_torch.py (ops and blocks), tests/test_clip_vs_hf.py (the
 * reference's clip.c on this oracle vs Hugging Face CLIP), tests/test_unet_vs_torch.py
 * (one full denoising step vs an LDM UNet in torch, SD1.x and SD2.x),
 * tests/test_vae_vs_torch.py (VAE decode / encode): agreement 3e-4 .. 1e-3.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load anything built from this file.
 *
 * Rounding points restated:
 *   mul_mat with F16 weights : activations rounded to f16, f32 accumulation
 *   conv_2d                  : im2col emitted in f16, f16 kernel, f32 accumulation
 *   norm / group_norm        : mean and variance accumulated in double
 *   soft_max                 : expf in f32, sum in double
 *   gelu / gelu_quick        : evaluated through an fp16 table (in and out f16)
 */
#define _GNU_SOURCE
#include "ggml.h"
#include "ggml-alloc.h"
#include "ggml-backend.h"
#include <math.h>
#include <stdarg.h>
#include <string.h>
#include <omp.h>

/* ------------------------------------------------------------------ */
/* fp16 / bf16 conversion (IEEE binary16, round to nearest even)       */

static inline float f16_to_f32(ggml_fp16_t h)
{ _Float16 x; memcpy(&x, &h, 2); return (float)x; }

static inline ggml_fp16_t f32_to_f16(float f)
{ _Float16 x = (_Float16)f; ggml_fp16_t h; memcpy(&h, &x, 2); return h; }

static inline float f16_round(float f) { return f16_to_f32(f32_to_f16(f)); }

float ggml_fp16_to_fp32(ggml_fp16_t x) { return f16_to_f32(x); }
ggml_fp16_t ggml_fp32_to_fp16(float x) { return f32_to_f16(x); }
void ggml_fp16_to_fp32_row(const ggml_fp16_t* x, float* y, int64_t n)
{ for (int64_t i = 0; i < n; ++i) y[i] = f16_to_f32(x[i]); }
void ggml_fp32_to_fp16_row(const float* x, ggml_fp16_t* y, int64_t n)
{ for (int64_t i = 0; i < n; ++i) y[i] = f32_to_f16(x[i]); }
void ggml_bf16_to_fp32_row(const ggml_bf16_t* x, float* y, int64_t n)
{
	for (int64_t i = 0; i < n; ++i) {
		uint32_t u = (uint32_t)x[i].bits << 16;
		memcpy(&y[i], &u, 4);
	}
}
size_t ggml_quantize_chunk(enum ggml_type type, const float* src, void* dst,
	int64_t start, int64_t nrows, int64_t n_per_row, const float* imatrix)
{
	(void)type; (void)src; (void)dst; (void)start; (void)nrows; (void)n_per_row; (void)imatrix;
	return 0;  /* quantised weights are outside the oracle's scope: caller reports an error */
}

void ggml_abort(const char* file, int line, const char* fmt, ...)
{
	va_list ap;
	fprintf(stderr, "[ggml_ref] %s:%d: ", file, line);
	va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap);
	fputc('\n', stderr);
	abort();
}

/* ------------------------------------------------------------------ */
/* types                                                               */

static const struct { const char* name; size_t size; } g_types[GGML_TYPE_COUNT] = {
	[GGML_TYPE_F32] = {"f32", 4}, [GGML_TYPE_F16] = {"f16", 2}, [GGML_TYPE_BF16] = {"bf16", 2},
	[GGML_TYPE_I8] = {"i8", 1}, [GGML_TYPE_I16] = {"i16", 2}, [GGML_TYPE_I32] = {"i32", 4},
	[GGML_TYPE_I64] = {"i64", 8}, [GGML_TYPE_F64] = {"f64", 8},
};

size_t ggml_type_size(enum ggml_type t)
{
	GGML_ASSERT((unsigned)t < GGML_TYPE_COUNT && g_types[t].size);
	return g_types[t].size;
}
const char* ggml_type_name(enum ggml_type t)
{
	return ((unsigned)t < GGML_TYPE_COUNT && g_types[t].name) ? g_types[t].name : "unsupported";
}
static void tt_f16_to_float(const void* x, float* y, int64_t k) { ggml_fp16_to_fp32_row(x, y, k); }
static void tt_bf16_to_float(const void* x, float* y, int64_t k) { ggml_bf16_to_fp32_row(x, y, k); }
const struct ggml_type_traits* ggml_get_type_traits(enum ggml_type t)
{
	static struct ggml_type_traits tr[GGML_TYPE_COUNT];
	GGML_ASSERT((unsigned)t < GGML_TYPE_COUNT);
	tr[t].type_name = ggml_type_name(t);
	tr[t].blck_size = 1;
	tr[t].type_size = g_types[t].size;
	tr[t].to_float = t == GGML_TYPE_F16 ? tt_f16_to_float : t == GGML_TYPE_BF16 ? tt_bf16_to_float : NULL;
	return &tr[t];
}

/* ------------------------------------------------------------------ */
/* context, tensors                                                    */

struct ggml_context {
	struct ggml_tensor** t;
	size_t n, cap;
	struct ggml_cgraph** graphs;
	size_t n_graphs;
};

struct ggml_cgraph {
	int size, n_nodes;
	struct ggml_tensor** nodes;
	struct ggml_tensor** seen;  /* visited list incl. leaves */
	int n_seen, cap_seen;
};

struct ggml_backend_buffer { int is_host; };
static struct ggml_backend_buffer g_host_buffer = { 1 };

struct ggml_context* ggml_init(struct ggml_init_params p)
{
	(void)p;
	return calloc(1, sizeof(struct ggml_context));
}

void ggml_free(struct ggml_context* ctx)
{
	if (!ctx) return;
	for (size_t i = 0; i < ctx->n; ++i) { free(ctx->t[i]->extra); free(ctx->t[i]); }
	for (size_t i = 0; i < ctx->n_graphs; ++i) {
		free(ctx->graphs[i]->nodes); free(ctx->graphs[i]->seen); free(ctx->graphs[i]);
	}
	free(ctx->t); free(ctx->graphs); free(ctx);
}

size_t ggml_tensor_overhead(void) { return sizeof(struct ggml_tensor) + 32; }
size_t ggml_graph_overhead(void) { return 1 << 16; }

static struct ggml_tensor* new_tensor(struct ggml_context* ctx, enum ggml_type type,
	int64_t n0, int64_t n1, int64_t n2, int64_t n3)
{
	struct ggml_tensor* t = calloc(1, sizeof(*t));
	t->type = type;
	t->ne[0] = n0; t->ne[1] = n1; t->ne[2] = n2; t->ne[3] = n3;
	t->nb[0] = ggml_type_size(type);
	for (int i = 1; i < 4; ++i) t->nb[i] = t->nb[i-1] * t->ne[i-1];
	if (ctx->n == ctx->cap) {
		ctx->cap = ctx->cap ? ctx->cap * 2 : 256;
		ctx->t = realloc(ctx->t, ctx->cap * sizeof(*ctx->t));
	}
	ctx->t[ctx->n++] = t;
	return t;
}

struct ggml_tensor* ggml_new_tensor_1d(struct ggml_context* c, enum ggml_type t, int64_t a)
{ return new_tensor(c, t, a, 1, 1, 1); }
struct ggml_tensor* ggml_new_tensor_2d(struct ggml_context* c, enum ggml_type t, int64_t a, int64_t b)
{ return new_tensor(c, t, a, b, 1, 1); }
struct ggml_tensor* ggml_new_tensor_4d(struct ggml_context* c, enum ggml_type t,
	int64_t a, int64_t b, int64_t d, int64_t e)
{ return new_tensor(c, t, a, b, d, e); }

struct ggml_tensor* ggml_get_first_tensor(const struct ggml_context* ctx)
{ return ctx->n ? ctx->t[0] : NULL; }
struct ggml_tensor* ggml_get_next_tensor(const struct ggml_context* ctx, struct ggml_tensor* t)
{
	for (size_t i = 0; i + 1 < ctx->n; ++i) if (ctx->t[i] == t) return ctx->t[i+1];
	return NULL;
}

struct ggml_tensor* ggml_set_name(struct ggml_tensor* t, const char* name)
{
	strncpy(t->name, name, sizeof(t->name) - 1);
	t->name[sizeof(t->name)-1] = 0;
	return t;
}
const char* ggml_get_name(const struct ggml_tensor* t) { return t->name; }
void ggml_set_input(struct ggml_tensor* t) { t->flags |= GGML_TENSOR_FLAG_INPUT; }
void ggml_set_output(struct ggml_tensor* t) { t->flags |= GGML_TENSOR_FLAG_OUTPUT; }
int64_t ggml_nelements(const struct ggml_tensor* t) { return t->ne[0]*t->ne[1]*t->ne[2]*t->ne[3]; }
size_t ggml_element_size(const struct ggml_tensor* t) { return ggml_type_size(t->type); }
size_t ggml_nbytes(const struct ggml_tensor* t)
{
	size_t n = ggml_type_size(t->type);
	for (int i = 0; i < 4; ++i) n += (t->ne[i] - 1) * t->nb[i];
	return n;
}
int ggml_n_dims(const struct ggml_tensor* t)
{
	for (int i = 3; i >= 1; --i) if (t->ne[i] > 1) return i + 1;
	return 1;
}

static const char* g_op_names[GGML_OP_COUNT] = {
	"NONE", "ADD", "MUL", "SCALE", "NORM", "GROUP_NORM", "MUL_MAT", "CONT", "RESHAPE", "VIEW",
	"PERMUTE", "TRANSPOSE", "GET_ROWS", "DIAG_MASK_INF", "SOFT_MAX", "CONV_2D", "CONCAT", "PAD",
	"UPSCALE", "TIMESTEP_EMBEDDING", "UNARY", "MAP_CUSTOM1",
};
static const char* g_unary_names[GGML_UNARY_OP_COUNT] = { "TANH", "RELU", "GELU", "GELU_QUICK", "SILU" };
const char* ggml_op_name(enum ggml_op op) { return (unsigned)op < GGML_OP_COUNT ? g_op_names[op] : "?"; }
const char* ggml_op_desc(const struct ggml_tensor* t)
{
	if (t->op == GGML_OP_UNARY) return g_unary_names[t->op_params[0]];
	return ggml_op_name(t->op);
}

/* ------------------------------------------------------------------ */
/* graph-op builders                                                    */

static bool is_contiguous(const struct ggml_tensor* t)
{
	size_t nb = ggml_type_size(t->type);
	for (int i = 0; i < 4; ++i) {
		if (t->ne[i] != 1 && t->nb[i] != nb) return false;
		nb *= t->ne[i];
	}
	return true;
}

static struct ggml_tensor* op_new(struct ggml_context* ctx, enum ggml_op op, enum ggml_type type,
	const int64_t ne[4], struct ggml_tensor* a, struct ggml_tensor* b)
{
	struct ggml_tensor* t = new_tensor(ctx, type, ne[0], ne[1], ne[2], ne[3]);
	t->op = op; t->src[0] = a; t->src[1] = b;
	return t;
}

/* A result that aliases a's storage (views, inplace ops). */
static struct ggml_tensor* op_view(struct ggml_context* ctx, enum ggml_op op,
	struct ggml_tensor* a, struct ggml_tensor* b)
{
	struct ggml_tensor* t = op_new(ctx, op, a->type, a->ne, a, b);
	memcpy(t->nb, a->nb, sizeof(t->nb));
	t->view_src = a->view_src ? a->view_src : a;
	t->view_offs = a->view_src ? a->view_offs : 0;
	return t;
}

static bool can_bcast(const struct ggml_tensor* a, const struct ggml_tensor* b)
{
	for (int i = 0; i < 4; ++i) if (a->ne[i] % b->ne[i]) return false;
	return true;
}

struct ggml_tensor* ggml_add(struct ggml_context* ctx, struct ggml_tensor* a, struct ggml_tensor* b)
{ GGML_ASSERT(can_bcast(a, b)); return op_new(ctx, GGML_OP_ADD, a->type, a->ne, a, b); }
struct ggml_tensor* ggml_add_inplace(struct ggml_context* ctx, struct ggml_tensor* a, struct ggml_tensor* b)
{ GGML_ASSERT(can_bcast(a, b)); return op_view(ctx, GGML_OP_ADD, a, b); }
struct ggml_tensor* ggml_mul(struct ggml_context* ctx, struct ggml_tensor* a, struct ggml_tensor* b)
{ GGML_ASSERT(can_bcast(a, b)); return op_new(ctx, GGML_OP_MUL, a->type, a->ne, a, b); }

static struct ggml_tensor* scale_impl(struct ggml_context* ctx, struct ggml_tensor* a, float s, bool inplace)
{
	struct ggml_tensor* t = inplace ? op_view(ctx, GGML_OP_SCALE, a, NULL)
	                                : op_new(ctx, GGML_OP_SCALE, a->type, a->ne, a, NULL);
	memcpy(t->op_params, &s, 4);
	return t;
}
struct ggml_tensor* ggml_scale(struct ggml_context* c, struct ggml_tensor* a, float s)
{ return scale_impl(c, a, s, false); }
struct ggml_tensor* ggml_scale_inplace(struct ggml_context* c, struct ggml_tensor* a, float s)
{ return scale_impl(c, a, s, true); }

struct ggml_tensor* ggml_mul_mat(struct ggml_context* ctx, struct ggml_tensor* a, struct ggml_tensor* b)
{
	GGML_ASSERT(a->ne[0] == b->ne[0] && b->ne[2] % a->ne[2] == 0 && b->ne[3] % a->ne[3] == 0);
	GGML_ASSERT(a->nb[0] == ggml_type_size(a->type));  /* a not transposed */
	int64_t ne[4] = { a->ne[1], b->ne[1], b->ne[2], b->ne[3] };
	return op_new(ctx, GGML_OP_MUL_MAT, GGML_TYPE_F32, ne, a, b);
}

struct ggml_tensor* ggml_conv_2d(struct ggml_context* ctx, struct ggml_tensor* a, struct ggml_tensor* b,
	int s0, int s1, int p0, int p1, int d0, int d1)
{
	GGML_ASSERT(a->ne[2] == b->ne[2]);
	int64_t ne[4] = {
		(b->ne[0] + 2*p0 - d0*(a->ne[0]-1) - 1) / s0 + 1,
		(b->ne[1] + 2*p1 - d1*(a->ne[1]-1) - 1) / s1 + 1,
		a->ne[3], b->ne[3] };
	struct ggml_tensor* t = op_new(ctx, GGML_OP_CONV_2D, GGML_TYPE_F32, ne, a, b);
	int32_t p[6] = { s0, s1, p0, p1, d0, d1 };
	memcpy(t->op_params, p, sizeof(p));
	return t;
}

struct ggml_tensor* ggml_norm(struct ggml_context* ctx, struct ggml_tensor* a, float eps)
{
	struct ggml_tensor* t = op_new(ctx, GGML_OP_NORM, a->type, a->ne, a, NULL);
	memcpy(t->op_params, &eps, 4);
	return t;
}
struct ggml_tensor* ggml_group_norm(struct ggml_context* ctx, struct ggml_tensor* a, int n_groups, float eps)
{
	struct ggml_tensor* t = op_new(ctx, GGML_OP_GROUP_NORM, a->type, a->ne, a, NULL);
	t->op_params[0] = n_groups;
	memcpy(&t->op_params[1], &eps, 4);
	return t;
}

static struct ggml_tensor* unary(struct ggml_context* ctx, struct ggml_tensor* a, int op, bool inplace)
{
	struct ggml_tensor* t = inplace ? op_view(ctx, GGML_OP_UNARY, a, NULL)
	                                : op_new(ctx, GGML_OP_UNARY, a->type, a->ne, a, NULL);
	t->op_params[0] = op;
	return t;
}
struct ggml_tensor* ggml_silu(struct ggml_context* c, struct ggml_tensor* a) { return unary(c, a, GGML_UNARY_OP_SILU, false); }
struct ggml_tensor* ggml_silu_inplace(struct ggml_context* c, struct ggml_tensor* a) { return unary(c, a, GGML_UNARY_OP_SILU, true); }
struct ggml_tensor* ggml_gelu_inplace(struct ggml_context* c, struct ggml_tensor* a) { return unary(c, a, GGML_UNARY_OP_GELU, true); }
struct ggml_tensor* ggml_gelu_quick_inplace(struct ggml_context* c, struct ggml_tensor* a) { return unary(c, a, GGML_UNARY_OP_GELU_QUICK, true); }
struct ggml_tensor* ggml_relu_inplace(struct ggml_context* c, struct ggml_tensor* a) { return unary(c, a, GGML_UNARY_OP_RELU, true); }
struct ggml_tensor* ggml_tanh_inplace(struct ggml_context* c, struct ggml_tensor* a) { return unary(c, a, GGML_UNARY_OP_TANH, true); }

struct ggml_tensor* ggml_soft_max_inplace(struct ggml_context* c, struct ggml_tensor* a)
{ return op_view(c, GGML_OP_SOFT_MAX, a, NULL); }
struct ggml_tensor* ggml_diag_mask_inf_inplace(struct ggml_context* c, struct ggml_tensor* a, int n_past)
{
	struct ggml_tensor* t = op_view(c, GGML_OP_DIAG_MASK_INF, a, NULL);
	t->op_params[0] = n_past;
	return t;
}

struct ggml_tensor* ggml_cont(struct ggml_context* c, struct ggml_tensor* a)
{ return op_new(c, GGML_OP_CONT, a->type, a->ne, a, NULL); }

struct ggml_tensor* ggml_permute(struct ggml_context* c, struct ggml_tensor* a, int x0, int x1, int x2, int x3)
{
	int ax[4] = { x0, x1, x2, x3 };
	struct ggml_tensor* t = op_view(c, GGML_OP_PERMUTE, a, NULL);
	for (int i = 0; i < 4; ++i) { t->ne[ax[i]] = a->ne[i]; t->nb[ax[i]] = a->nb[i]; }
	memcpy(t->op_params, ax, sizeof(ax));
	return t;
}
struct ggml_tensor* ggml_transpose(struct ggml_context* c, struct ggml_tensor* a)
{
	struct ggml_tensor* t = op_view(c, GGML_OP_TRANSPOSE, a, NULL);
	t->ne[0] = a->ne[1]; t->ne[1] = a->ne[0];
	t->nb[0] = a->nb[1]; t->nb[1] = a->nb[0];
	return t;
}

static struct ggml_tensor* reshape(struct ggml_context* c, struct ggml_tensor* a,
	int64_t n0, int64_t n1, int64_t n2, int64_t n3)
{
	GGML_ASSERT(is_contiguous(a));
	GGML_ASSERT(ggml_nelements(a) == n0*n1*n2*n3);
	struct ggml_tensor* t = op_view(c, GGML_OP_RESHAPE, a, NULL);
	t->ne[0] = n0; t->ne[1] = n1; t->ne[2] = n2; t->ne[3] = n3;
	t->nb[0] = ggml_type_size(a->type);
	for (int i = 1; i < 4; ++i) t->nb[i] = t->nb[i-1] * t->ne[i-1];
	return t;
}
struct ggml_tensor* ggml_reshape_3d(struct ggml_context* c, struct ggml_tensor* a, int64_t n0, int64_t n1, int64_t n2)
{ return reshape(c, a, n0, n1, n2, 1); }
struct ggml_tensor* ggml_reshape_4d(struct ggml_context* c, struct ggml_tensor* a, int64_t n0, int64_t n1, int64_t n2, int64_t n3)
{ return reshape(c, a, n0, n1, n2, n3); }

struct ggml_tensor* ggml_view_4d(struct ggml_context* c, struct ggml_tensor* a,
	int64_t n0, int64_t n1, int64_t n2, int64_t n3, size_t nb1, size_t nb2, size_t nb3, size_t offset)
{
	struct ggml_tensor* t = op_view(c, GGML_OP_VIEW, a, NULL);
	t->ne[0] = n0; t->ne[1] = n1; t->ne[2] = n2; t->ne[3] = n3;
	t->nb[0] = ggml_type_size(a->type); t->nb[1] = nb1; t->nb[2] = nb2; t->nb[3] = nb3;
	t->view_offs += offset;
	memcpy(t->op_params, &offset, sizeof(offset));
	return t;
}
struct ggml_tensor* ggml_view_1d(struct ggml_context* c, struct ggml_tensor* a, int64_t n0, size_t offset)
{
	size_t es = ggml_type_size(a->type);
	return ggml_view_4d(c, a, n0, 1, 1, 1, es*n0, es*n0, es*n0, offset);
}

struct ggml_tensor* ggml_concat(struct ggml_context* c, struct ggml_tensor* a, struct ggml_tensor* b, int dim)
{
	int64_t ne[4];
	for (int i = 0; i < 4; ++i) {
		if (i == dim) ne[i] = a->ne[i] + b->ne[i];
		else { GGML_ASSERT(a->ne[i] == b->ne[i]); ne[i] = a->ne[i]; }
	}
	struct ggml_tensor* t = op_new(c, GGML_OP_CONCAT, a->type, ne, a, b);
	t->op_params[0] = dim;
	return t;
}
struct ggml_tensor* ggml_pad(struct ggml_context* c, struct ggml_tensor* a, int p0, int p1, int p2, int p3)
{
	int64_t ne[4] = { a->ne[0]+p0, a->ne[1]+p1, a->ne[2]+p2, a->ne[3]+p3 };
	return op_new(c, GGML_OP_PAD, a->type, ne, a, NULL);
}
struct ggml_tensor* ggml_upscale(struct ggml_context* c, struct ggml_tensor* a, int sf, enum ggml_scale_mode mode)
{
	GGML_ASSERT(mode == GGML_SCALE_MODE_NEAREST);
	int64_t ne[4] = { a->ne[0]*sf, a->ne[1]*sf, a->ne[2], a->ne[3] };
	struct ggml_tensor* t = op_new(c, GGML_OP_UPSCALE, a->type, ne, a, NULL);
	t->op_params[0] = mode;
	return t;
}
struct ggml_tensor* ggml_get_rows(struct ggml_context* c, struct ggml_tensor* a, struct ggml_tensor* b)
{
	GGML_ASSERT(a->ne[2] == b->ne[1] && b->ne[3] == 1 && b->type == GGML_TYPE_I32);
	int64_t ne[4] = { a->ne[0], b->ne[0], b->ne[1], b->ne[2] };
	return op_new(c, GGML_OP_GET_ROWS, GGML_TYPE_F32, ne, a, b);
}
struct ggml_tensor* ggml_timestep_embedding(struct ggml_context* c, struct ggml_tensor* ts, int dim, int max_period)
{
	int64_t ne[4] = { dim + (dim & 1), ts->ne[0], 1, 1 };
	struct ggml_tensor* t = op_new(c, GGML_OP_TIMESTEP_EMBEDDING, GGML_TYPE_F32, ne, ts, NULL);
	t->op_params[0] = dim; t->op_params[1] = max_period;
	return t;
}
struct ggml_tensor* ggml_map_custom1_inplace(struct ggml_context* c, struct ggml_tensor* a,
	ggml_custom1_op_t fun, int n_tasks, void* userdata)
{
	struct ggml_tensor* t = op_view(c, GGML_OP_MAP_CUSTOM1, a, NULL);
	struct { ggml_custom1_op_t fun; int n_tasks; void* ud; } p = { fun, n_tasks, userdata };
	memcpy(t->op_params, &p, sizeof(p));
	return t;
}

/* ------------------------------------------------------------------ */
/* graph                                                                */

struct ggml_cgraph* ggml_new_graph_custom(struct ggml_context* ctx, size_t size, bool grads)
{
	(void)grads;
	struct ggml_cgraph* g = calloc(1, sizeof(*g));
	g->size = (int)size;
	g->nodes = calloc(size, sizeof(*g->nodes));
	ctx->graphs = realloc(ctx->graphs, (ctx->n_graphs + 1) * sizeof(*ctx->graphs));
	ctx->graphs[ctx->n_graphs++] = g;
	return g;
}

static void visit(struct ggml_cgraph* g, struct ggml_tensor* t)
{
	/* "seen" marker kept in the tensor's private padding (one graph per context pair) */
	if (t->padding[0] == 1) return;
	t->padding[0] = 1;
	if (g->n_seen == g->cap_seen) {
		g->cap_seen = g->cap_seen ? g->cap_seen * 2 : 1024;
		g->seen = realloc(g->seen, g->cap_seen * sizeof(*g->seen));
	}
	g->seen[g->n_seen++] = t;
	for (int i = 0; i < GGML_MAX_SRC; ++i) if (t->src[i]) visit(g, t->src[i]);
	if (t->op != GGML_OP_NONE) {
		GGML_ASSERT(g->n_nodes < g->size);
		g->nodes[g->n_nodes++] = t;
	}
}
void ggml_build_forward_expand(struct ggml_cgraph* g, struct ggml_tensor* t) { visit(g, t); }
int ggml_graph_size(struct ggml_cgraph* g) { return g->size; }
int ggml_graph_n_nodes(struct ggml_cgraph* g) { return g->n_nodes; }

/* ------------------------------------------------------------------ */
/* allocator: leaves and OUTPUT tensors get permanent storage here;     */
/* intermediates are allocated and released during compute.             */

struct ggml_gallocr { void** bufs; size_t n, total; struct ggml_cgraph* graph; };

ggml_gallocr_t ggml_gallocr_new(ggml_backend_buffer_type_t buft)
{ (void)buft; return calloc(1, sizeof(struct ggml_gallocr)); }

static void galloc_release(ggml_gallocr_t a)
{
	for (size_t i = 0; i < a->n; ++i) free(a->bufs[i]);
	free(a->bufs); a->bufs = NULL; a->n = 0; a->total = 0;
}
void ggml_gallocr_free(ggml_gallocr_t a) { if (a) { galloc_release(a); free(a); } }

static bool is_alias_op(const struct ggml_tensor* t) { return t->view_src != NULL; }

static void* galloc_buf(ggml_gallocr_t a, size_t sz)
{
	void* p = calloc(1, sz + 64);
	a->bufs = realloc(a->bufs, (a->n + 1) * sizeof(void*));
	a->bufs[a->n++] = p;
	a->total += sz;
	return p;
}

bool ggml_gallocr_reserve(ggml_gallocr_t a, struct ggml_cgraph* g)
{
	galloc_release(a);
	a->graph = g;
	for (int i = 0; i < g->n_seen; ++i) {  /* an OUTPUT alias keeps its storage root alive */
		struct ggml_tensor* t = g->seen[i];
		if (is_alias_op(t) && (t->flags & GGML_TENSOR_FLAG_OUTPUT)) t->view_src->flags |= GGML_TENSOR_FLAG_OUTPUT;
	}
	for (int i = 0; i < g->n_seen; ++i) {
		struct ggml_tensor* t = g->seen[i];
		if (is_alias_op(t)) continue;
		if (t->op == GGML_OP_NONE || (t->flags & GGML_TENSOR_FLAG_OUTPUT)) {
			t->data = galloc_buf(a, ggml_nbytes(t));
			t->buffer = &g_host_buffer;
		}
	}
	return true;
}
bool ggml_gallocr_alloc_graph(ggml_gallocr_t a, struct ggml_cgraph* g)
{
	if (a->graph != g) return ggml_gallocr_reserve(a, g);
	return true;
}
size_t ggml_gallocr_get_buffer_size(ggml_gallocr_t a, int id) { (void)id; return a->total; }

/* ------------------------------------------------------------------ */
/* backend registry (a single "CPU" reference backend)                  */

struct ggml_backend { int n_threads; };
struct ggml_backend_reg { int dummy; };
struct ggml_backend_device { int dummy; };
struct ggml_backend_buffer_type { int dummy; };
static struct ggml_backend_reg g_reg;
static struct ggml_backend_device g_dev;
static struct ggml_backend_buffer_type g_buft;

ggml_backend_t ggml_backend_init_by_name(const char* name, const char* params)
{
	(void)name; (void)params;
	struct ggml_backend* b = calloc(1, sizeof(*b));
	b->n_threads = omp_get_max_threads();
	return b;
}
ggml_backend_t ggml_backend_init_best(void) { return ggml_backend_init_by_name("CPU", NULL); }
void ggml_backend_free(ggml_backend_t b) { free(b); }
const char* ggml_backend_name(ggml_backend_t b) { (void)b; return "CPU-ref-oracle"; }
ggml_backend_buffer_type_t ggml_backend_get_default_buffer_type(ggml_backend_t b) { (void)b; return &g_buft; }
ggml_backend_dev_t ggml_backend_get_device(ggml_backend_t b) { (void)b; return &g_dev; }
bool ggml_backend_buffer_is_host(ggml_backend_buffer_t b) { return b ? b->is_host : true; }
size_t ggml_backend_reg_count(void) { return 1; }
ggml_backend_reg_t ggml_backend_reg_get(size_t i) { return i == 0 ? &g_reg : NULL; }
const char* ggml_backend_reg_name(ggml_backend_reg_t r) { (void)r; return "CPU"; }
size_t ggml_backend_reg_dev_count(ggml_backend_reg_t r) { (void)r; return 1; }
ggml_backend_dev_t ggml_backend_reg_dev_get(ggml_backend_reg_t r, size_t i) { (void)r; return i == 0 ? &g_dev : NULL; }
static void set_n_threads(ggml_backend_t b, int n) { b->n_threads = n; omp_set_num_threads(n); }
void* ggml_backend_reg_get_proc_address(ggml_backend_reg_t r, const char* name)
{
	(void)r;
	if (!strcmp(name, "ggml_backend_set_n_threads")) return (void*)set_n_threads;
	return NULL;
}
const char* ggml_backend_dev_name(ggml_backend_dev_t d) { (void)d; return "CPU"; }
const char* ggml_backend_dev_description(ggml_backend_dev_t d) { (void)d; return "reference-semantics CPU oracle"; }
void ggml_backend_dev_memory(ggml_backend_dev_t d, size_t* f, size_t* t) { (void)d; *f = 0; *t = 0; }
ggml_backend_reg_t ggml_backend_dev_backend_reg(ggml_backend_dev_t d) { (void)d; return &g_reg; }

static void* tensor_ptr(const struct ggml_tensor* t)
{
	if (t->view_src) return (char*)tensor_ptr(t->view_src) + t->view_offs;
	return t->data;
}

void ggml_backend_tensor_set(struct ggml_tensor* t, const void* data, size_t off, size_t size)
{
	GGML_ASSERT(tensor_ptr(t) && off + size <= ggml_nbytes(t));
	memcpy((char*)tensor_ptr(t) + off, data, size);
	struct ggml_tensor* root = t->view_src ? t->view_src : t;
	free(root->extra); root->extra = NULL;  /* drop the cached f32 copy */
}
void ggml_backend_tensor_get(const struct ggml_tensor* t, void* data, size_t off, size_t size)
{
	GGML_ASSERT(tensor_ptr(t) && off + size <= ggml_nbytes(t));
	memcpy(data, (const char*)tensor_ptr(t) + off, size);
}

/* ------------------------------------------------------------------ */
/* op execution                                                         */

#define NE(t)  const int64_t t##0 = (t)->ne[0], t##1 = (t)->ne[1], t##2 = (t)->ne[2], t##3 = (t)->ne[3]
#define AT(t, p, i0, i1, i2, i3) \
	((char*)(p) + (i0)*(t)->nb[0] + (i1)*(t)->nb[1] + (i2)*(t)->nb[2] + (i3)*(t)->nb[3])

static inline float ld(const struct ggml_tensor* t, const void* p, int64_t i0, int64_t i1, int64_t i2, int64_t i3)
{
	const char* q = AT(t, p, i0, i1, i2, i3);
	return t->type == GGML_TYPE_F16 ? f16_to_f32(*(const ggml_fp16_t*)q) : *(const float*)q;
}
static inline void st(const struct ggml_tensor* t, void* p, int64_t i0, int64_t i1, int64_t i2, int64_t i3, float v)
{
	char* q = AT(t, p, i0, i1, i2, i3);
	if (t->type == GGML_TYPE_F16) *(ggml_fp16_t*)q = f32_to_f16(v); else *(float*)q = v;
}

/* ggml_add / ggml_mul: src1 broadcast by modulo; result has src0's type
 * (F16 for lora.c:63 add_inplace: one rounding per merge). */
static void op_binary(struct ggml_tensor* d, int is_mul)
{
	const struct ggml_tensor *a = d->src[0], *b = d->src[1];
	const void *pa = tensor_ptr(a), *pb = tensor_ptr(b); void* pd = tensor_ptr(d);
	NE(d);
	#pragma omp parallel for collapse(2) schedule(static)
	for (int64_t i3 = 0; i3 < d3; ++i3)
	for (int64_t i2 = 0; i2 < d2; ++i2)
	for (int64_t i1 = 0; i1 < d1; ++i1)
	for (int64_t i0 = 0; i0 < d0; ++i0) {
		float x = ld(a, pa, i0, i1, i2, i3);
		float y = ld(b, pb, i0 % b->ne[0], i1 % b->ne[1], i2 % b->ne[2], i3 % b->ne[3]);
		st(d, pd, i0, i1, i2, i3, is_mul ? x * y : x + y);
	}
}

static void op_scale(struct ggml_tensor* d)
{
	const struct ggml_tensor* a = d->src[0];
	const void* pa = tensor_ptr(a); void* pd = tensor_ptr(d);
	float s; memcpy(&s, d->op_params, 4);
	NE(d);
	#pragma omp parallel for collapse(2) schedule(static)
	for (int64_t i3 = 0; i3 < d3; ++i3)
	for (int64_t i2 = 0; i2 < d2; ++i2)
	for (int64_t i1 = 0; i1 < d1; ++i1)
	for (int64_t i0 = 0; i0 < d0; ++i0)
		st(d, pd, i0, i1, i2, i3, ld(a, pa, i0, i1, i2, i3) * s);
}

static void op_cont(struct ggml_tensor* d)
{
	const struct ggml_tensor* a = d->src[0];
	const void* pa = tensor_ptr(a); void* pd = tensor_ptr(d);
	NE(d);
	#pragma omp parallel for collapse(2) schedule(static)
	for (int64_t i3 = 0; i3 < d3; ++i3)
	for (int64_t i2 = 0; i2 < d2; ++i2)
	for (int64_t i1 = 0; i1 < d1; ++i1)
	for (int64_t i0 = 0; i0 < d0; ++i0)
		st(d, pd, i0, i1, i2, i3, ld(a, pa, i0, i1, i2, i3));
}

/* ggml_norm (mlblock_nn.c:65): per row over ne[0], biased variance, double sums. */
static void op_norm(struct ggml_tensor* d)
{
	const struct ggml_tensor* a = d->src[0];
	const void* pa = tensor_ptr(a); void* pd = tensor_ptr(d);
	float eps; memcpy(&eps, d->op_params, 4);
	NE(d);
	#pragma omp parallel for collapse(3) schedule(static)
	for (int64_t i3 = 0; i3 < d3; ++i3)
	for (int64_t i2 = 0; i2 < d2; ++i2)
	for (int64_t i1 = 0; i1 < d1; ++i1) {
		double sum = 0;
		for (int64_t i0 = 0; i0 < d0; ++i0) sum += (double)ld(a, pa, i0, i1, i2, i3);
		float mean = (float)(sum / d0);
		double sum2 = 0;
		for (int64_t i0 = 0; i0 < d0; ++i0) {
			float v = ld(a, pa, i0, i1, i2, i3) - mean;
			st(d, pd, i0, i1, i2, i3, v);
			sum2 += (double)(v * v);
		}
		float variance = (float)(sum2 / d0);
		const float scale = 1.0f / sqrtf(variance + eps);
		for (int64_t i0 = 0; i0 < d0; ++i0)
			st(d, pd, i0, i1, i2, i3, ld(d, pd, i0, i1, i2, i3) * scale);
	}
}

/* ggml_group_norm (mlblock_nn.c:86): groups of ceil(C/G) channels x W x H, per N. */
static void op_group_norm(struct ggml_tensor* d)
{
	const struct ggml_tensor* a = d->src[0];
	const void* pa = tensor_ptr(a); void* pd = tensor_ptr(d);
	int n_groups = d->op_params[0];
	float eps; memcpy(&eps, &d->op_params[1], 4);
	NE(d);
	int64_t cpg = (d2 + n_groups - 1) / n_groups;
	#pragma omp parallel for collapse(2) schedule(static)
	for (int64_t i3 = 0; i3 < d3; ++i3)
	for (int64_t g = 0; g < n_groups; ++g) {
		int64_t c0 = g * cpg, c1 = c0 + cpg > d2 ? d2 : c0 + cpg;
		if (c0 >= c1) continue;
		double sum = 0;
		for (int64_t i2 = c0; i2 < c1; ++i2)
		for (int64_t i1 = 0; i1 < d1; ++i1) {
			double rs = 0;
			for (int64_t i0 = 0; i0 < d0; ++i0) rs += (double)ld(a, pa, i0, i1, i2, i3);
			sum += rs;
		}
		const float mean = (float)(sum / (d0 * d1 * (c1 - c0)));
		double sum2 = 0;
		for (int64_t i2 = c0; i2 < c1; ++i2)
		for (int64_t i1 = 0; i1 < d1; ++i1) {
			double rs = 0;
			for (int64_t i0 = 0; i0 < d0; ++i0) {
				float v = ld(a, pa, i0, i1, i2, i3) - mean;
				st(d, pd, i0, i1, i2, i3, v);
				rs += (double)(v * v);
			}
			sum2 += rs;
		}
		const float variance = (float)(sum2 / (d0 * d1 * (c1 - c0)));
		const float scale = 1.0f / sqrtf(variance + eps);
		for (int64_t i2 = c0; i2 < c1; ++i2)
		for (int64_t i1 = 0; i1 < d1; ++i1)
		for (int64_t i0 = 0; i0 < d0; ++i0)
			st(d, pd, i0, i1, i2, i3, ld(d, pd, i0, i1, i2, i3) * scale);
	}
}

static inline float gelu_f32(float x)
{ return 0.5f * x * (1.0f + tanhf(0.79788456080286535587989211986876f * x * (1.0f + 0.044715f * x * x))); }
static inline float gelu_quick_f32(float x) { return x * (1.0f / (1.0f + expf(-1.702f * x))); }

static void op_unary(struct ggml_tensor* d)
{
	const struct ggml_tensor* a = d->src[0];
	const void* pa = tensor_ptr(a); void* pd = tensor_ptr(d);
	int op = d->op_params[0];
	NE(d);
	#pragma omp parallel for collapse(2) schedule(static)
	for (int64_t i3 = 0; i3 < d3; ++i3)
	for (int64_t i2 = 0; i2 < d2; ++i2)
	for (int64_t i1 = 0; i1 < d1; ++i1)
	for (int64_t i0 = 0; i0 < d0; ++i0) {
		float x = ld(a, pa, i0, i1, i2, i3), y;
		switch (op) {
		case GGML_UNARY_OP_TANH: y = tanhf(x); break;
		case GGML_UNARY_OP_RELU: y = x > 0 ? x : 0; break;
		case GGML_UNARY_OP_SILU: y = x / (1.0f + expf(-x)); break;
		case GGML_UNARY_OP_GELU:  /* fp16 lookup-table semantics, |x|>=10 clamped */
			if (x <= -10.0f) y = 0.0f; else if (x >= 10.0f) y = x;
			else y = f16_round(gelu_f32(f16_round(x)));
			break;
		case GGML_UNARY_OP_GELU_QUICK:
			y = f16_round(gelu_quick_f32(f16_round(x)));
			break;
		default: GGML_ABORT("unary op %d", op);
		}
		st(d, pd, i0, i1, i2, i3, y);
	}
}

/* ggml_soft_max over ne[0] (ggml_extend.c:217). */
static void op_soft_max(struct ggml_tensor* d)
{
	const struct ggml_tensor* a = d->src[0];
	const void* pa = tensor_ptr(a); void* pd = tensor_ptr(d);
	NE(d);
	#pragma omp parallel for collapse(3) schedule(static)
	for (int64_t i3 = 0; i3 < d3; ++i3)
	for (int64_t i2 = 0; i2 < d2; ++i2)
	for (int64_t i1 = 0; i1 < d1; ++i1) {
		float mx = -INFINITY;
		for (int64_t i0 = 0; i0 < d0; ++i0) { float v = ld(a, pa, i0, i1, i2, i3); if (v > mx) mx = v; }
		double sum = 0;
		for (int64_t i0 = 0; i0 < d0; ++i0) {
			float e = expf(ld(a, pa, i0, i1, i2, i3) - mx);
			st(d, pd, i0, i1, i2, i3, e);
			sum += (double)e;
		}
		float inv = (float)(1.0 / sum);
		for (int64_t i0 = 0; i0 < d0; ++i0) st(d, pd, i0, i1, i2, i3, ld(d, pd, i0, i1, i2, i3) * inv);
	}
}

/* ggml_diag_mask_inf(x, n_past) (ggml_extend.c:215): -inf where key i0 > n_past + query i1. */
static void op_diag_mask_inf(struct ggml_tensor* d)
{
	const struct ggml_tensor* a = d->src[0];
	const void* pa = tensor_ptr(a); void* pd = tensor_ptr(d);
	int n_past = d->op_params[0];
	NE(d);
	for (int64_t i3 = 0; i3 < d3; ++i3)
	for (int64_t i2 = 0; i2 < d2; ++i2)
	for (int64_t i1 = 0; i1 < d1; ++i1)
	for (int64_t i0 = 0; i0 < d0; ++i0)
		st(d, pd, i0, i1, i2, i3, i0 > n_past + i1 ? -INFINITY : ld(a, pa, i0, i1, i2, i3));
}

/* Row-times-rowT GEMM: C[n][m] = sum_k B[n][k] * A[m][k], f32 accumulation.
 * Both operands K-contiguous. Blocked 4x4 with the k loop vectorised. */
static void gemm_nt(int64_t M, int64_t N, int64_t K,
	const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc)
{
	#pragma omp parallel for schedule(dynamic, 1)
	for (int64_t n0 = 0; n0 < N; n0 += 4) {
		int64_t nn = N - n0 < 4 ? N - n0 : 4;
		for (int64_t m0 = 0; m0 < M; m0 += 4) {
			int64_t mm = M - m0 < 4 ? M - m0 : 4;
			if (nn == 4 && mm == 4) {
				const float *b0 = B + (n0+0)*ldb, *b1 = B + (n0+1)*ldb, *b2 = B + (n0+2)*ldb, *b3 = B + (n0+3)*ldb;
				const float *a0 = A + (m0+0)*lda, *a1 = A + (m0+1)*lda, *a2 = A + (m0+2)*lda, *a3 = A + (m0+3)*lda;
				float c00=0,c01=0,c02=0,c03=0,c10=0,c11=0,c12=0,c13=0,c20=0,c21=0,c22=0,c23=0,c30=0,c31=0,c32=0,c33=0;
				#pragma omp simd reduction(+:c00,c01,c02,c03,c10,c11,c12,c13,c20,c21,c22,c23,c30,c31,c32,c33)
				for (int64_t k = 0; k < K; ++k) {
					float x0 = b0[k], x1 = b1[k], x2 = b2[k], x3 = b3[k];
					float y0 = a0[k], y1 = a1[k], y2 = a2[k], y3 = a3[k];
					c00 += x0*y0; c01 += x0*y1; c02 += x0*y2; c03 += x0*y3;
					c10 += x1*y0; c11 += x1*y1; c12 += x1*y2; c13 += x1*y3;
					c20 += x2*y0; c21 += x2*y1; c22 += x2*y2; c23 += x2*y3;
					c30 += x3*y0; c31 += x3*y1; c32 += x3*y2; c33 += x3*y3;
				}
				float* c = C + n0*ldc + m0;
				c[0]=c00; c[1]=c01; c[2]=c02; c[3]=c03; c += ldc;
				c[0]=c10; c[1]=c11; c[2]=c12; c[3]=c13; c += ldc;
				c[0]=c20; c[1]=c21; c[2]=c22; c[3]=c23; c += ldc;
				c[0]=c30; c[1]=c31; c[2]=c32; c[3]=c33;
			} else {
				for (int64_t n = n0; n < n0 + nn; ++n)
				for (int64_t m = m0; m < m0 + mm; ++m) {
					const float *b = B + n*ldb, *a = A + m*lda;
					float acc = 0;
					#pragma omp simd reduction(+:acc)
					for (int64_t k = 0; k < K; ++k) acc += b[k] * a[k];
					C[n*ldc + m] = acc;
				}
			}
		}
	}
}

/* Gather rows [nrows][K] of a strided tensor plane into dense f32, optionally rounding to f16. */
static float* rows_to_f32(const struct ggml_tensor* t, const void* p, int64_t i2, int64_t i3, bool round16)
{
	int64_t K = t->ne[0], R = t->ne[1];
	float* out = malloc((size_t)K * R * sizeof(float) + 64);
	GGML_ASSERT(out);
	#pragma omp parallel for schedule(static)
	for (int64_t r = 0; r < R; ++r)
		for (int64_t k = 0; k < K; ++k) {
			float v = ld(t, p, k, r, i2, i3);
			out[r*K + k] = round16 ? f16_round(v) : v;
		}
	return out;
}

/* ggml_mul_mat(a,b) (mlblock_nn.c:22, ggml_extend.c:212,219, clip.c:433, lora.c:60):
 * a:[K,M,a2,a3] b:[K,N,b2,b3] -> [M,N,b2,b3]; b is converted to a's dot type first:
 * F16 weights => activations rounded to f16; products accumulated in f32. */
static void op_mul_mat(struct ggml_tensor* d)
{
	const struct ggml_tensor *a = d->src[0], *b = d->src[1];
	const void *pa = tensor_ptr(a), *pb = tensor_ptr(b); void* pd = tensor_ptr(d);
	GGML_ASSERT(d->type == GGML_TYPE_F32 && d->nb[0] == 4 && d->nb[1] == (size_t)d->ne[0]*4);
	bool round16 = a->type == GGML_TYPE_F16;
	int64_t r2 = b->ne[2] / a->ne[2], r3 = b->ne[3] / a->ne[3];
	float* af = NULL; int64_t af2 = -1, af3 = -1;
	/* weights (2-d leaves) keep a cached f32 copy across computes */
	bool a_leaf = a->op == GGML_OP_NONE && !a->view_src && a->ne[2] == 1 && a->ne[3] == 1;
	for (int64_t i3 = 0; i3 < b->ne[3]; ++i3)
	for (int64_t i2 = 0; i2 < b->ne[2]; ++i2) {
		int64_t j2 = i2 / r2, j3 = i3 / r3;
		if (j2 != af2 || j3 != af3) {
			if (!a_leaf) free(af);
			if (a_leaf && a->extra) af = a->extra;
			else {
				af = rows_to_f32(a, pa, j2, j3, false);
				if (a_leaf) ((struct ggml_tensor*)a)->extra = af;
			}
			af2 = j2; af3 = j3;
		}
		float* bf = rows_to_f32(b, pb, i2, i3, round16);
		gemm_nt(a->ne[1], b->ne[1], a->ne[0], af, a->ne[0], bf, b->ne[0],
			(float*)AT(d, pd, 0, 0, i2, i3), d->ne[0]);
		free(bf);
	}
	if (!a_leaf) free(af);
}

/* ggml_conv_2d(w,x,...) (mlblock_nn.c:44): im2col in the kernel's type (F16), then mul_mat;
 * w:[KW,KH,Cin,Cout] x:[W,H,Cin,N] -> [OW,OH,Cout,N], zero padding. */
static void op_conv_2d(struct ggml_tensor* d)
{
	const struct ggml_tensor *w = d->src[0], *x = d->src[1];
	const void *pw = tensor_ptr(w), *px = tensor_ptr(x); float* pd = tensor_ptr(d);
	const int32_t* p = d->op_params;
	int s0 = p[0], s1 = p[1], p0 = p[2], p1 = p[3], d0 = p[4], d1 = p[5];
	int64_t KW = w->ne[0], KH = w->ne[1], C = w->ne[2], OC = w->ne[3];
	int64_t W = x->ne[0], H = x->ne[1], N = x->ne[3], OW = d->ne[0], OH = d->ne[1];
	int64_t K = KW * KH * C;
	bool round16 = w->type == GGML_TYPE_F16;
	GGML_ASSERT(is_contiguous(w) && is_contiguous(d));
	bool w_leaf = w->op == GGML_OP_NONE && !w->view_src;
	float* wf = w_leaf ? w->extra : NULL;
	if (!wf) {
		wf = malloc((size_t)K * OC * sizeof(float));
		#pragma omp parallel for schedule(static)
		for (int64_t i = 0; i < K * OC; ++i)
			wf[i] = w->type == GGML_TYPE_F16 ? f16_to_f32(((const ggml_fp16_t*)pw)[i]) : ((const float*)pw)[i];
		if (w_leaf) ((struct ggml_tensor*)w)->extra = wf;
	}
	float* col = malloc((size_t)OW * OH * K * sizeof(float) + 64);
	float* out = malloc((size_t)OW * OH * OC * sizeof(float) + 64);
	GGML_ASSERT(wf && col && out);
	for (int64_t n = 0; n < N; ++n) {
		#pragma omp parallel for collapse(2) schedule(static)
		for (int64_t oh = 0; oh < OH; ++oh)
		for (int64_t ow = 0; ow < OW; ++ow) {
			float* c = col + (oh*OW + ow) * K;
			for (int64_t ci = 0; ci < C; ++ci)
			for (int64_t kh = 0; kh < KH; ++kh)
			for (int64_t kw = 0; kw < KW; ++kw) {
				int64_t iw = ow*s0 + kw*d0 - p0, ih = oh*s1 + kh*d1 - p1;
				float v = 0;
				if (iw >= 0 && iw < W && ih >= 0 && ih < H) {
					v = ld(x, px, iw, ih, ci, n);
					if (round16) v = f16_round(v);
				}
				c[(ci*KH + kh)*KW + kw] = v;
			}
		}
		gemm_nt(OC, OW*OH, K, wf, K, col, K, out, OC);
		#pragma omp parallel for schedule(static)
		for (int64_t oc = 0; oc < OC; ++oc)
			for (int64_t i = 0; i < OW*OH; ++i)
				pd[(n*OC + oc) * OW*OH + i] = out[i*OC + oc];
	}
	if (!w_leaf) free(wf);
	free(col); free(out);
}

static void op_concat(struct ggml_tensor* d)
{
	const struct ggml_tensor *a = d->src[0], *b = d->src[1];
	const void *pa = tensor_ptr(a), *pb = tensor_ptr(b); void* pd = tensor_ptr(d);
	int dim = d->op_params[0];
	NE(d);
	#pragma omp parallel for collapse(2) schedule(static)
	for (int64_t i3 = 0; i3 < d3; ++i3)
	for (int64_t i2 = 0; i2 < d2; ++i2)
	for (int64_t i1 = 0; i1 < d1; ++i1)
	for (int64_t i0 = 0; i0 < d0; ++i0) {
		int64_t i[4] = { i0, i1, i2, i3 };
		float v;
		if (i[dim] < a->ne[dim]) v = ld(a, pa, i0, i1, i2, i3);
		else { i[dim] -= a->ne[dim]; v = ld(b, pb, i[0], i[1], i[2], i[3]); }
		st(d, pd, i0, i1, i2, i3, v);
	}
}

/* ggml_pad: zero-pad at the END of each dim (mlblock_nn.c:110). */
static void op_pad(struct ggml_tensor* d)
{
	const struct ggml_tensor* a = d->src[0];
	const void* pa = tensor_ptr(a); void* pd = tensor_ptr(d);
	NE(d);
	#pragma omp parallel for collapse(2) schedule(static)
	for (int64_t i3 = 0; i3 < d3; ++i3)
	for (int64_t i2 = 0; i2 < d2; ++i2)
	for (int64_t i1 = 0; i1 < d1; ++i1)
	for (int64_t i0 = 0; i0 < d0; ++i0) {
		bool in = i0 < a->ne[0] && i1 < a->ne[1] && i2 < a->ne[2] && i3 < a->ne[3];
		st(d, pd, i0, i1, i2, i3, in ? ld(a, pa, i0, i1, i2, i3) : 0.0f);
	}
}

/* ggml_upscale NEAREST: out[i0,i1] = x[i0/sf, i1/sf] (mlblock_nn.c:122, tae.c:82). */
static void op_upscale(struct ggml_tensor* d)
{
	const struct ggml_tensor* a = d->src[0];
	const void* pa = tensor_ptr(a); void* pd = tensor_ptr(d);
	NE(d);
	int64_t f0 = d0 / a->ne[0], f1 = d1 / a->ne[1];
	#pragma omp parallel for collapse(2) schedule(static)
	for (int64_t i3 = 0; i3 < d3; ++i3)
	for (int64_t i2 = 0; i2 < d2; ++i2)
	for (int64_t i1 = 0; i1 < d1; ++i1)
	for (int64_t i0 = 0; i0 < d0; ++i0)
		st(d, pd, i0, i1, i2, i3, ld(a, pa, i0 / f0, i1 / f1, i2, i3));
}

/* ggml_get_rows(table, ids) (clip.c:338): F16/F32 table -> f32 rows. */
static void op_get_rows(struct ggml_tensor* d)
{
	const struct ggml_tensor *a = d->src[0], *b = d->src[1];
	const void *pa = tensor_ptr(a), *pb = tensor_ptr(b); void* pd = tensor_ptr(d);
	NE(d);
	for (int64_t i3 = 0; i3 < d3; ++i3)
	for (int64_t i2 = 0; i2 < d2; ++i2)
	for (int64_t i1 = 0; i1 < d1; ++i1) {
		int32_t row = *(const int32_t*)AT(b, pb, i1, i2, i3, 0);
		GGML_ASSERT(row >= 0 && row < a->ne[1]);
		for (int64_t i0 = 0; i0 < d0; ++i0)
			st(d, pd, i0, i1, i2, i3, ld(a, pa, i0, row, i2, 0));
	}
}

/* ggml_timestep_embedding (unet.c:150): cos first, then sin. */
static void op_timestep_embedding(struct ggml_tensor* d)
{
	const struct ggml_tensor* a = d->src[0];
	const float* ts = tensor_ptr(a); float* pd = tensor_ptr(d);
	int dim = d->op_params[0], max_period = d->op_params[1];
	int half = dim / 2;
	for (int64_t i = 0; i < a->ne[0]; ++i) {
		float* e = (float*)((char*)pd + i * d->nb[1]);
		for (int j = 0; j < half; ++j) {
			float timestep = ts[i];
			float freq = expf(-logf((float)max_period) * j / half);
			float arg = timestep * freq;
			e[j] = cosf(arg);
			e[j + half] = sinf(arg);
		}
		if (dim & 1) e[dim] = 0.0f;
	}
}

static void op_map_custom1(struct ggml_tensor* d)
{
	struct { ggml_custom1_op_t fun; int n_tasks; void* ud; } p;
	memcpy(&p, d->op_params, sizeof(p));
	/* debug taps read dst->data / src->data directly (ggml_extend.c:161-166) */
	struct ggml_tensor dd = *d, ss = *d->src[0];
	dd.data = tensor_ptr(d); ss.data = tensor_ptr(d->src[0]);
	p.fun(&dd, &ss, 0, 1, p.ud);
}

/* Test knob: GGML_REF_ROUND=f16 rounds every materialised node to f16, emulating an
 * engine that stores activations in half precision (used to size the error budget). */
static int g_round_mode = -1;
static void maybe_round(struct ggml_tensor* t)
{
	if (g_round_mode < 0) {
		const char* e = getenv("GGML_REF_ROUND");
		g_round_mode = (e && !strcmp(e, "f16")) ? 1 : 0;
	}
	if (!g_round_mode || t->type != GGML_TYPE_F32 || !is_contiguous(t)) return;
	float* p = tensor_ptr(t);
	int64_t n = ggml_nelements(t);
	#pragma omp parallel for schedule(static)
	for (int64_t i = 0; i < n; ++i) p[i] = f16_round(p[i]);
}

enum ggml_status ggml_backend_graph_compute(ggml_backend_t backend, struct ggml_cgraph* g)
{
	(void)backend;
	int n = g->n_nodes;
	/* liveness: last consumer index of each storage root */
	int* last = malloc(n * sizeof(int));
	for (int i = 0; i < n; ++i) last[i] = -1;
	/* index nodes through padding[4..7] */
	for (int i = 0; i < n; ++i) memcpy(g->nodes[i]->padding + 4, &i, 4);
	for (int i = 0; i < n; ++i) {
		struct ggml_tensor* t = g->nodes[i];
		for (int s = 0; s < GGML_MAX_SRC; ++s) {
			struct ggml_tensor* r = t->src[s];
			if (!r) continue;
			if (r->view_src) r = r->view_src;
			if (r->op == GGML_OP_NONE) continue;
			int idx; memcpy(&idx, r->padding + 4, 4);
			if (idx >= 0 && idx < n && g->nodes[idx] == r) last[idx] = i;
		}
	}
	char* owned = calloc(n, 1);
	for (int i = 0; i < n; ++i) {
		struct ggml_tensor* t = g->nodes[i];
		if (!t->view_src && !(t->flags & GGML_TENSOR_FLAG_OUTPUT)) {
			t->data = malloc(ggml_nbytes(t) + 64);
			GGML_ASSERT(t->data);
			t->buffer = &g_host_buffer;
			owned[i] = 1;
		}
		switch (t->op) {
		case GGML_OP_ADD: op_binary(t, 0); break;
		case GGML_OP_MUL: op_binary(t, 1); break;
		case GGML_OP_SCALE: op_scale(t); break;
		case GGML_OP_NORM: op_norm(t); break;
		case GGML_OP_GROUP_NORM: op_group_norm(t); break;
		case GGML_OP_MUL_MAT: op_mul_mat(t); break;
		case GGML_OP_CONV_2D: op_conv_2d(t); break;
		case GGML_OP_CONT: op_cont(t); break;
		case GGML_OP_RESHAPE: case GGML_OP_VIEW: case GGML_OP_PERMUTE: case GGML_OP_TRANSPOSE: break;
		case GGML_OP_GET_ROWS: op_get_rows(t); break;
		case GGML_OP_DIAG_MASK_INF: op_diag_mask_inf(t); break;
		case GGML_OP_SOFT_MAX: op_soft_max(t); break;
		case GGML_OP_CONCAT: op_concat(t); break;
		case GGML_OP_PAD: op_pad(t); break;
		case GGML_OP_UPSCALE: op_upscale(t); break;
		case GGML_OP_TIMESTEP_EMBEDDING: op_timestep_embedding(t); break;
		case GGML_OP_UNARY: op_unary(t); break;
		case GGML_OP_MAP_CUSTOM1: op_map_custom1(t); break;
		default: GGML_ABORT("op %d not implemented", (int)t->op);
		}
		if (!t->view_src || t->op == GGML_OP_ADD || t->op == GGML_OP_SCALE || t->op == GGML_OP_UNARY ||
			t->op == GGML_OP_SOFT_MAX)
			maybe_round(t);
		/* release storage whose last consumer was this node */
		for (int j = 0; j <= i; ++j)
			if (owned[j] && last[j] <= i && (last[j] == i || (j == i && last[j] < 0))) {
				free(g->nodes[j]->data); g->nodes[j]->data = NULL; owned[j] = 0;
			}
	}
	for (int i = 0; i < n; ++i) if (owned[i]) { free(g->nodes[i]->data); g->nodes[i]->data = NULL; }
	free(owned); free(last);
	return GGML_STATUS_SUCCESS;
}
