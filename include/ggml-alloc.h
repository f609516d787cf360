/* ggml-alloc.h -- graph allocator half of the drop-in boundary.
 * Call sites: mlblock.c:161-180 (mlctx_alloc), mlblock.c:22-25 (mlctx_free).
 * In the B200 engine "allocation" gives every leaf (parameter / input) and
 * every tensor flagged OUTPUT a device buffer in the logical ggml layout;
 * intermediate activations live in a planner-owned arena instead.
 */
#ifndef GGML_B200_GGML_ALLOC_H
#define GGML_B200_GGML_ALLOC_H
#include "ggml.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct ggml_backend_buffer_type* ggml_backend_buffer_type_t;
typedef struct ggml_backend_buffer*      ggml_backend_buffer_t;
typedef struct ggml_backend*             ggml_backend_t;
typedef struct ggml_gallocr*             ggml_gallocr_t;

GGML_API ggml_gallocr_t ggml_gallocr_new(ggml_backend_buffer_type_t buft);
GGML_API void   ggml_gallocr_free(ggml_gallocr_t galloc);
GGML_API bool   ggml_gallocr_reserve(ggml_gallocr_t galloc, struct ggml_cgraph* graph);
GGML_API bool   ggml_gallocr_alloc_graph(ggml_gallocr_t galloc, struct ggml_cgraph* graph);
GGML_API size_t ggml_gallocr_get_buffer_size(ggml_gallocr_t galloc, int buffer_id);

#ifdef __cplusplus
}
#endif
#endif
