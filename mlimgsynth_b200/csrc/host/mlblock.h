/* mlblock.h -- module-graph interface of the B200 host layer.
 *
 * Same contract as the reference's mlblock.h:44-160 (MLCtx; mlctx_begin / block_begin /
 * tensor_add / input_new / prep / compute / end), so the model builders (unet.c, vae.c, tae.c,
 * clip.c) read like the reference's: a block function builds ops through the ggml-shaped ABI and
 * the caller names the result afterwards with mlctx_tensor_add(); parameter paths are
 * reconstructed from the nesting when the graph is prepared.
 * Differences that matter on B200: a prepared MLCtx (graph + device-resident weights + captured
 * CUDA graph) is kept alive and reused across generations instead of being rebuilt and re-uploaded
 * for every call (the reference reloads every graph: clip.c:460-486, mlimgsynth.c:1733,1753).
 */
#pragma once
#include "base.h"
#include "tstore.h"
#include "ggml.h"
#include "ggml-alloc.h"
#include "ggml-backend.h"
#include "ggml-b200.h"

typedef struct ggml_tensor MLTensor;

enum { MLB_F_QUIET = 2 };

typedef struct MLCtxEntry { MLTensor* t; char* name; int kind; char* key; } MLCtxEntry;  /* kind: 0 named tensor, 1 block begin */

typedef struct MLCtx {
	ggml_backend_t backend;      /* fill before use */
	TStore* tstore;              /* fill before use */
	struct ggml_context *cp, *cc;
	struct ggml_cgraph* graph;
	ggml_gallocr_t allocr;
	MLCtxEntry* ent; int n_ent, cap_ent;
	MLTensor** inputs; int n_inputs, cap_inputs;
	MLTensor* result;
	struct { enum ggml_type wtype; unsigned n_tensor_max; const char* tprefix; const char* name; int flags; } c;
	struct { size_t mem_params, mem_total; double t_load, t_compute; unsigned n_compute, n_conv; } info;
	bool prepared;
} MLCtx;

void mlctx_begin(MLCtx* C, const char* name);
void mlctx_end(MLCtx* C);                      /* frees graph, contexts and device memory */
void mlctx_block_begin(MLCtx* C);
MLTensor* mlctx_tensor_add(MLCtx* C, const char* name, MLTensor* tensor);
MLTensor* mlctx_input_new(MLCtx* C, const char* name, enum ggml_type dtype, int n0, int n1, int n2, int n3);
int mlctx_prep(MLCtx* C);                      /* resolve names, build, allocate, upload parameters */
int mlctx_compute(MLCtx* C);                   /* asynchronous on the engine stream */
int mlctx_reload_params(MLCtx* C);             /* re-upload parameters (after a LoRA merge changed the store) */

/* host <-> graph tensors (localtensor.h:96-106 role); MLIS-style host tensor */
typedef struct HTensor { float* d; int n[4]; int flags; } HTensor;
enum { HT_OWNMEM = 1, HT_READY = 2 };
void   ht_resize(HTensor* T, int n0, int n1, int n2, int n3);
void   ht_free(HTensor* T);
void   ht_copy(HTensor* dst, const HTensor* src);
size_t ht_count(const HTensor* T);
int    ht_finite_check(const HTensor* T);
