/* ggml-backend.h -- backend half of the drop-in boundary.
 * Call sites: mlimgsynth.c:572-598 (backend enumeration), :1120-1165 (init,
 * thread count), mlblock.c:294-307 (graph compute), localtensor.h:96-106 and
 * unet.c:375-384 (tensor upload / download each NFE), ggml_extend.c:176,195.
 *
 * One backend is registered: "B200" (aliases "CUDA", "CUDA0", "GPU"); device i
 * of the process is reachable as "B200:<i>" / "CUDA<i>". There is no CPU
 * backend and no fallback: with no usable sm_100 device init returns NULL.
 */
#ifndef GGML_B200_GGML_BACKEND_H
#define GGML_B200_GGML_BACKEND_H
#include "ggml.h"
#include "ggml-alloc.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct ggml_backend_reg*    ggml_backend_reg_t;
typedef struct ggml_backend_device* ggml_backend_dev_t;

enum ggml_status {
	GGML_STATUS_ALLOC_FAILED = -2,
	GGML_STATUS_FAILED       = -1,
	GGML_STATUS_SUCCESS      = 0,
	GGML_STATUS_ABORTED      = 1,
};

typedef void (*ggml_backend_set_n_threads_t)(ggml_backend_t backend, int n_threads);

/* backend lifecycle */
GGML_API ggml_backend_t ggml_backend_init_by_name(const char* name, const char* params);
GGML_API ggml_backend_t ggml_backend_init_best(void);
GGML_API void           ggml_backend_free(ggml_backend_t backend);
GGML_API const char*    ggml_backend_name(ggml_backend_t backend);
GGML_API ggml_backend_buffer_type_t ggml_backend_get_default_buffer_type(ggml_backend_t backend);
GGML_API ggml_backend_dev_t ggml_backend_get_device(ggml_backend_t backend);

/* Runs the recorded graph. Returns 0 on success (mlblock.c:301-307). */
GGML_API enum ggml_status ggml_backend_graph_compute(ggml_backend_t backend,
	struct ggml_cgraph* cgraph);

/* Synchronous byte copies between host memory and the tensor's logical layout. */
GGML_API void ggml_backend_tensor_set(struct ggml_tensor* tensor,
	const void* data, size_t offset, size_t size);
GGML_API void ggml_backend_tensor_get(const struct ggml_tensor* tensor,
	void* data, size_t offset, size_t size);
GGML_API bool ggml_backend_buffer_is_host(ggml_backend_buffer_t buffer);

/* registry / device enumeration */
GGML_API size_t             ggml_backend_reg_count(void);
GGML_API ggml_backend_reg_t ggml_backend_reg_get(size_t index);
GGML_API const char*        ggml_backend_reg_name(ggml_backend_reg_t reg);
GGML_API size_t             ggml_backend_reg_dev_count(ggml_backend_reg_t reg);
GGML_API ggml_backend_dev_t ggml_backend_reg_dev_get(ggml_backend_reg_t reg, size_t index);
GGML_API void*              ggml_backend_reg_get_proc_address(ggml_backend_reg_t reg, const char* name);
GGML_API const char*        ggml_backend_dev_name(ggml_backend_dev_t device);
GGML_API const char*        ggml_backend_dev_description(ggml_backend_dev_t device);
GGML_API void               ggml_backend_dev_memory(ggml_backend_dev_t device, size_t* free, size_t* total);
GGML_API ggml_backend_reg_t ggml_backend_dev_backend_reg(ggml_backend_dev_t device);

#ifdef __cplusplus
}
#endif
#endif
