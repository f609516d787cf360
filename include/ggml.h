/* ggml.h -- the graph-builder half of the drop-in boundary.
 *
 * This header declares the ggml-shaped C ABI that libmlimgsynth's model code
 * (unet.c, vae.c, tae.c, clip.c, lora.c, mlblock.c, mlblock_nn.c,
 * ggml_extend.c, localtensor.h, tensorstore.c) compiles against. It is NOT
 * ggml: there is no CPU executor behind it. Every builder only records a node;
 * the work happens in ggml_backend_graph_compute() (ggml-backend.h), which
 * plans the recorded graph into fused sm_100a CUDA kernels.
 *
 * Two libraries export these symbols:
 *   - mlimgsynth_b200/csrc/  -> libggml_b200.so   (the product, CUDA sm_100a)
 *   - oracle/ggml_ref.c      -> test-only CPU restatement of the op semantics
 *
 * Each declaration cites the reference call site it serves (paths relative to
 * the reference tree, src/...). Layout of struct ggml_tensor follows upstream
 * ggml so that objects compiled against upstream headers stay ABI-compatible.
 */
#ifndef GGML_B200_GGML_H
#define GGML_B200_GGML_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GGML_API

#define GGML_MAX_DIMS            4      /* ggml_extend.c:138 */
#define GGML_MAX_OP_PARAMS      64
#define GGML_MAX_SRC            10
#define GGML_MAX_NAME           64      /* ggml_extend.c:16-29 name prefixing */
#define GGML_DEFAULT_GRAPH_SIZE 2048    /* mlblock.c:57 */

#define GGML_ABORT(...) ggml_abort(__FILE__, __LINE__, __VA_ARGS__)
#define GGML_ASSERT(x) \
	do { if (!(x)) GGML_ABORT("GGML_ASSERT(%s) failed", #x); } while (0)

GGML_API void ggml_abort(const char* file, int line, const char* fmt, ...);

typedef uint16_t ggml_fp16_t;
typedef struct { uint16_t bits; } ggml_bf16_t;

/* Values are the on-disk / upstream ones; tensorstore.c:83-96 maps to them. */
enum ggml_type {
	GGML_TYPE_F32     = 0,
	GGML_TYPE_F16     = 1,
	GGML_TYPE_Q4_0    = 2,
	GGML_TYPE_Q4_1    = 3,
	GGML_TYPE_Q5_0    = 6,
	GGML_TYPE_Q5_1    = 7,
	GGML_TYPE_Q8_0    = 8,
	GGML_TYPE_Q8_1    = 9,
	GGML_TYPE_Q2_K    = 10,
	GGML_TYPE_Q3_K    = 11,
	GGML_TYPE_Q4_K    = 12,
	GGML_TYPE_Q5_K    = 13,
	GGML_TYPE_Q6_K    = 14,
	GGML_TYPE_Q8_K    = 15,
	GGML_TYPE_IQ2_XXS = 16,
	GGML_TYPE_IQ2_XS  = 17,
	GGML_TYPE_IQ3_XXS = 18,
	GGML_TYPE_IQ1_S   = 19,
	GGML_TYPE_IQ4_NL  = 20,
	GGML_TYPE_IQ3_S   = 21,
	GGML_TYPE_IQ2_S   = 22,
	GGML_TYPE_IQ4_XS  = 23,
	GGML_TYPE_I8      = 24,
	GGML_TYPE_I16     = 25,
	GGML_TYPE_I32     = 26,
	GGML_TYPE_I64     = 27,
	GGML_TYPE_F64     = 28,
	GGML_TYPE_IQ1_M   = 29,
	GGML_TYPE_BF16    = 30,
	GGML_TYPE_TQ1_0   = 34,
	GGML_TYPE_TQ2_0   = 35,
	GGML_TYPE_COUNT   = 39,
};

/* Only GGML_OP_NONE == 0 is relied upon by callers (mlblock.h:126,
 * mlblock.c:90,125,275: "op == NONE" means parameter/input leaf). */
enum ggml_op {
	GGML_OP_NONE = 0,
	GGML_OP_ADD,
	GGML_OP_MUL,
	GGML_OP_SCALE,
	GGML_OP_NORM,
	GGML_OP_GROUP_NORM,
	GGML_OP_MUL_MAT,
	GGML_OP_CONT,
	GGML_OP_RESHAPE,
	GGML_OP_VIEW,
	GGML_OP_PERMUTE,
	GGML_OP_TRANSPOSE,
	GGML_OP_GET_ROWS,
	GGML_OP_DIAG_MASK_INF,
	GGML_OP_SOFT_MAX,
	GGML_OP_CONV_2D,
	GGML_OP_CONCAT,
	GGML_OP_PAD,
	GGML_OP_UPSCALE,
	GGML_OP_TIMESTEP_EMBEDDING,
	GGML_OP_UNARY,
	GGML_OP_MAP_CUSTOM1,
	GGML_OP_COUNT,
};

enum ggml_unary_op {
	GGML_UNARY_OP_TANH = 0,
	GGML_UNARY_OP_RELU,
	GGML_UNARY_OP_GELU,
	GGML_UNARY_OP_GELU_QUICK,
	GGML_UNARY_OP_SILU,
	GGML_UNARY_OP_COUNT,
};

enum ggml_scale_mode {
	GGML_SCALE_MODE_NEAREST  = 0,   /* mlblock_nn.c:122, tae.c:82 */
	GGML_SCALE_MODE_BILINEAR = 1,
};

enum ggml_tensor_flag {
	GGML_TENSOR_FLAG_INPUT  = 1,
	GGML_TENSOR_FLAG_OUTPUT = 2,
	GGML_TENSOR_FLAG_PARAM  = 4,
};

struct ggml_context;
struct ggml_cgraph;
struct ggml_backend_buffer;

/* Read directly by callers: ne/nb (mlblock_nn.c:20,38,63,196-198, unet.c:118,
 * clip.c:337-339,431, ggml_extend.c:141-152), op, name, type, data, buffer. */
struct ggml_tensor {
	enum ggml_type type;
	struct ggml_backend_buffer* buffer;
	int64_t ne[GGML_MAX_DIMS];   /* elements, ne[0] innermost */
	size_t  nb[GGML_MAX_DIMS];   /* byte strides of the LOGICAL (ggml) layout */
	enum ggml_op op;
	int32_t op_params[GGML_MAX_OP_PARAMS / sizeof(int32_t)];
	int32_t flags;
	struct ggml_tensor* src[GGML_MAX_SRC];
	struct ggml_tensor* view_src;
	size_t view_offs;
	void* data;                  /* device address of the logical-layout copy, if any */
	char name[GGML_MAX_NAME];
	void* extra;                 /* engine-private per-tensor record */
	char padding[8];
};

struct ggml_init_params {        /* mlblock.c:60 */
	size_t mem_size;
	void*  mem_buffer;
	bool   no_alloc;
};

typedef void (*ggml_to_float_t)(const void* x, float* y, int64_t k);
typedef void (*ggml_from_float_t)(const float* x, void* y, int64_t k);

struct ggml_type_traits {        /* tensorstore.c:222-224 (dequantise on host) */
	const char* type_name;
	int64_t blck_size;
	int64_t blck_size_interleave;
	size_t  type_size;
	bool    is_quantized;
	ggml_to_float_t   to_float;
	ggml_from_float_t from_float_ref;
};

typedef void (*ggml_custom1_op_t)(struct ggml_tensor* dst,
	const struct ggml_tensor* a, int ith, int nth, void* userdata);

/* ---- context / graph lifecycle (mlblock.c:54-63, 20-47, 107-150) ---- */
GGML_API struct ggml_context* ggml_init(struct ggml_init_params params);
GGML_API void   ggml_free(struct ggml_context* ctx);
GGML_API size_t ggml_tensor_overhead(void);
GGML_API size_t ggml_graph_overhead(void);

GGML_API struct ggml_tensor* ggml_new_tensor_1d(struct ggml_context* ctx,
	enum ggml_type type, int64_t ne0);
GGML_API struct ggml_tensor* ggml_new_tensor_2d(struct ggml_context* ctx,
	enum ggml_type type, int64_t ne0, int64_t ne1);
GGML_API struct ggml_tensor* ggml_new_tensor_4d(struct ggml_context* ctx,
	enum ggml_type type, int64_t ne0, int64_t ne1, int64_t ne2, int64_t ne3);

GGML_API struct ggml_cgraph* ggml_new_graph_custom(struct ggml_context* ctx,
	size_t size, bool grads);
GGML_API void ggml_build_forward_expand(struct ggml_cgraph* cgraph,
	struct ggml_tensor* tensor);
GGML_API int  ggml_graph_size(struct ggml_cgraph* cgraph);
GGML_API int  ggml_graph_n_nodes(struct ggml_cgraph* cgraph);

GGML_API struct ggml_tensor* ggml_get_first_tensor(const struct ggml_context* ctx);
GGML_API struct ggml_tensor* ggml_get_next_tensor(const struct ggml_context* ctx,
	struct ggml_tensor* tensor);

/* ---- tensor metadata ---- */
GGML_API struct ggml_tensor* ggml_set_name(struct ggml_tensor* tensor, const char* name);
GGML_API const char* ggml_get_name(const struct ggml_tensor* tensor);
GGML_API void ggml_set_input(struct ggml_tensor* tensor);   /* mlblock.h:147,157 */
GGML_API void ggml_set_output(struct ggml_tensor* tensor);  /* mlblock.c:133,140, unet.c:413-414 */
GGML_API size_t  ggml_nbytes(const struct ggml_tensor* tensor);
GGML_API int64_t ggml_nelements(const struct ggml_tensor* tensor);
GGML_API size_t  ggml_element_size(const struct ggml_tensor* tensor);
GGML_API size_t  ggml_type_size(enum ggml_type type);
GGML_API const char* ggml_type_name(enum ggml_type type);
GGML_API int ggml_n_dims(const struct ggml_tensor* tensor);
GGML_API const char* ggml_op_name(enum ggml_op op);
GGML_API const char* ggml_op_desc(const struct ggml_tensor* t);
GGML_API const struct ggml_type_traits* ggml_get_type_traits(enum ggml_type type);

/* ---- graph-op builders: record one node, return it ---- */
/* elementwise, 2nd operand broadcast (mlblock_nn.c:25,51,68,72,99,101,144,154) */
GGML_API struct ggml_tensor* ggml_add(struct ggml_context* ctx,
	struct ggml_tensor* a, struct ggml_tensor* b);
GGML_API struct ggml_tensor* ggml_add_inplace(struct ggml_context* ctx,
	struct ggml_tensor* a, struct ggml_tensor* b);                 /* lora.c:63 */
GGML_API struct ggml_tensor* ggml_mul(struct ggml_context* ctx,
	struct ggml_tensor* a, struct ggml_tensor* b);
GGML_API struct ggml_tensor* ggml_scale(struct ggml_context* ctx,
	struct ggml_tensor* a, float s);                               /* vae.c:174, tae.c:71,73 */
GGML_API struct ggml_tensor* ggml_scale_inplace(struct ggml_context* ctx,
	struct ggml_tensor* a, float s);                               /* ggml_extend.c:213, lora.c:62 */

/* contractions */
GGML_API struct ggml_tensor* ggml_mul_mat(struct ggml_context* ctx,
	struct ggml_tensor* a, struct ggml_tensor* b);                 /* mlblock_nn.c:22 */
GGML_API struct ggml_tensor* ggml_conv_2d(struct ggml_context* ctx,
	struct ggml_tensor* a, struct ggml_tensor* b,
	int s0, int s1, int p0, int p1, int d0, int d1);               /* mlblock_nn.c:44 */

/* normalisation */
GGML_API struct ggml_tensor* ggml_norm(struct ggml_context* ctx,
	struct ggml_tensor* a, float eps);                             /* mlblock_nn.c:65 */
GGML_API struct ggml_tensor* ggml_group_norm(struct ggml_context* ctx,
	struct ggml_tensor* a, int n_groups, float eps);               /* mlblock_nn.c:86 */

/* activations */
GGML_API struct ggml_tensor* ggml_silu(struct ggml_context* ctx, struct ggml_tensor* a);
GGML_API struct ggml_tensor* ggml_silu_inplace(struct ggml_context* ctx, struct ggml_tensor* a);
GGML_API struct ggml_tensor* ggml_gelu_inplace(struct ggml_context* ctx, struct ggml_tensor* a);
GGML_API struct ggml_tensor* ggml_gelu_quick_inplace(struct ggml_context* ctx, struct ggml_tensor* a);
GGML_API struct ggml_tensor* ggml_relu_inplace(struct ggml_context* ctx, struct ggml_tensor* a);
GGML_API struct ggml_tensor* ggml_tanh_inplace(struct ggml_context* ctx, struct ggml_tensor* a);

/* attention pieces (ggml_extend.c:200-221) */
GGML_API struct ggml_tensor* ggml_soft_max_inplace(struct ggml_context* ctx,
	struct ggml_tensor* a);
GGML_API struct ggml_tensor* ggml_diag_mask_inf_inplace(struct ggml_context* ctx,
	struct ggml_tensor* a, int n_past);

/* layout */
GGML_API struct ggml_tensor* ggml_cont(struct ggml_context* ctx, struct ggml_tensor* a);
GGML_API struct ggml_tensor* ggml_permute(struct ggml_context* ctx,
	struct ggml_tensor* a, int axis0, int axis1, int axis2, int axis3);
GGML_API struct ggml_tensor* ggml_transpose(struct ggml_context* ctx, struct ggml_tensor* a);
GGML_API struct ggml_tensor* ggml_reshape_3d(struct ggml_context* ctx,
	struct ggml_tensor* a, int64_t ne0, int64_t ne1, int64_t ne2);
GGML_API struct ggml_tensor* ggml_reshape_4d(struct ggml_context* ctx,
	struct ggml_tensor* a, int64_t ne0, int64_t ne1, int64_t ne2, int64_t ne3);
GGML_API struct ggml_tensor* ggml_view_1d(struct ggml_context* ctx,
	struct ggml_tensor* a, int64_t ne0, size_t offset);            /* clip.c:431 */
GGML_API struct ggml_tensor* ggml_view_4d(struct ggml_context* ctx,
	struct ggml_tensor* a, int64_t ne0, int64_t ne1, int64_t ne2, int64_t ne3,
	size_t nb1, size_t nb2, size_t nb3, size_t offset);            /* ggml_extend.c:152 */
GGML_API struct ggml_tensor* ggml_concat(struct ggml_context* ctx,
	struct ggml_tensor* a, struct ggml_tensor* b, int dim);        /* unet.c:233 */
GGML_API struct ggml_tensor* ggml_pad(struct ggml_context* ctx,
	struct ggml_tensor* a, int p0, int p1, int p2, int p3);        /* mlblock_nn.c:110 */
GGML_API struct ggml_tensor* ggml_upscale(struct ggml_context* ctx,
	struct ggml_tensor* a, int scale_factor, enum ggml_scale_mode mode);

/* misc */
GGML_API struct ggml_tensor* ggml_get_rows(struct ggml_context* ctx,
	struct ggml_tensor* a, struct ggml_tensor* b);                 /* clip.c:338 */
GGML_API struct ggml_tensor* ggml_timestep_embedding(struct ggml_context* ctx,
	struct ggml_tensor* timesteps, int dim, int max_period);       /* unet.c:150 */
GGML_API struct ggml_tensor* ggml_map_custom1_inplace(struct ggml_context* ctx,
	struct ggml_tensor* a, ggml_custom1_op_t fun, int n_tasks, void* userdata);

/* ---- host conversion helpers (tensorstore.c:191-225, lora.c:83) ---- */
GGML_API float ggml_fp16_to_fp32(ggml_fp16_t x);
GGML_API ggml_fp16_t ggml_fp32_to_fp16(float x);
GGML_API void ggml_fp16_to_fp32_row(const ggml_fp16_t* x, float* y, int64_t n);
GGML_API void ggml_fp32_to_fp16_row(const float* x, ggml_fp16_t* y, int64_t n);
GGML_API void ggml_bf16_to_fp32_row(const ggml_bf16_t* x, float* y, int64_t n);
GGML_API size_t ggml_quantize_chunk(enum ggml_type type, const float* src,
	void* dst, int64_t start, int64_t nrows, int64_t n_per_row, const float* imatrix);

#ifdef __cplusplus
}
#endif
#endif
