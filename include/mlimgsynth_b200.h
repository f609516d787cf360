/* mlimgsynth_b200.h -- public C API of libmlimgsynth_b200.so.
 *
 * ABI-compatible with the reference's include/mlimgsynth.h (v0.4.2): same function names,
 * argument meaning, enum values, struct layouts and error convention, so a program or binding
 * written against the reference header (python/mlimgsynth.py, main_mlimgsynth.c) can load this
 * library instead. Differences, all additive:
 *   - MLIS_OPT_BATCH_SIZE > 1 is implemented (the reference rejects it, mlimgsynth.c:1640):
 *     image i of a batch is generated with seed + i and a fresh noise offset, i.e. what the
 *     reference's generate.sh:55-61 loop produces; mlis_image_get(ctx, i) returns image i and
 *     MLIS_TENSOR_LATENT / MLIS_TENSOR_IMAGE carry the batch in n[3];
 *   - the only backend is "B200" (CUDA sm_100a); MLIS_OPT_THREADS and MLIS_OPT_UNET_SPLIT are
 *     accepted and ignored (no host compute threads; 180 GB of HBM needs no graph splitting);
 *   - weights stay resident on the device between generations.
 */
#ifndef MLIMGSYNTH_B200_H
#define MLIMGSYNTH_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MLIS_VERSION      0x000402
#define MLIS_VERSION_STR  "0.4.2"

typedef enum MLIS_ErrorCode {
	MLIS_E_UNKNOWN = -1, MLIS_E_VERSION = -2, MLIS_E_UNK_OPT = -3, MLIS_E_OPT_VALUE = -4,
	MLIS_E_PROMPT_PARSE = -5, MLIS_E_FILE_NOT_FOUND = -6, MLIS_E_NAN = -7, MLIS_E_IMAGE = -8,
} MLIS_ErrorCode;

typedef enum MLIS_Stage {
	MLIS_STAGE_IDLE = 0, MLIS_STAGE_COND_ENCODE = 1, MLIS_STAGE_IMAGE_ENCODE = 2,
	MLIS_STAGE_IMAGE_DECODE = 3, MLIS_STAGE_DENOISE = 4,
} MLIS_Stage;

typedef enum MLIS_Method {
	MLIS_METHOD_NONE = 0, MLIS_METHOD_EULER = 1, MLIS_METHOD_HEUN = 2, MLIS_METHOD_TAYLOR3 = 3,
	MLIS_METHOD_DPMPP2M = 4, MLIS_METHOD_DPMPP2S = 5, MLIS_METHOD__LAST = 5,
} MLIS_Method;

typedef enum MLIS_Scheduler { MLIS_SCHED_NONE = 0, MLIS_SCHED_UNIFORM = 1, MLIS_SCHED_KARRAS = 2, MLIS_SCHED__LAST = 2 } MLIS_Scheduler;

typedef enum MLIS_LogLvl {
	MLIS_LOGLVL_NONE = 0, MLIS_LOGLVL_ERROR = 10, MLIS_LOGLVL_WARNING = 20, MLIS_LOGLVL_INFO = 30,
	MLIS_LOGLVL_VERBOSE = 40, MLIS_LOGLVL_DEBUG = 50, MLIS_LOGLVL_MAX = 255,
	MLIS_LOGLVL__INCREASE = 0x100 | 10, MLIS_LOGLVL__DECREASE = 0x200 | 10,
} MLIS_LogLvl;

typedef enum MLIS_TensorId {
	MLIS_TENSOR_IMAGE = 1, MLIS_TENSOR_MASK = 2, MLIS_TENSOR_LATENT = 3, MLIS_TENSOR_LMASK = 4,
	MLIS_TENSOR_COND = 5, MLIS_TENSOR_LABEL = 6, MLIS_TENSOR_NCOND = 7, MLIS_TENSOR_NLABEL = 8,
	MLIS_TENSOR_TMP = 0x100,
} MLIS_TensorId;

typedef enum MLIS_TensorUseFlag {
	MLIS_TUF_IMAGE = 1, MLIS_TUF_MASK = 2, MLIS_TUF_LATENT = 4, MLIS_TUF_LMASK = 8, MLIS_TUF_CONDITIONING = 16,
} MLIS_TensorUseFlag;

typedef enum MLIS_ModelType {
	MLIS_MODEL_TYPE_NONE = 0, MLIS_MODEL_TYPE_SD1 = 1, MLIS_MODEL_TYPE_SD2 = 2, MLIS_MODEL_TYPE_SDXL = 3, MLIS_MODEL_TYPE__LAST = 3,
} MLIS_ModelType;

typedef enum MLIS_SubModel {
	MLIS_SUBMODEL_NONE = 0, MLIS_SUBMODEL_UNET = 1, MLIS_SUBMODEL_VAE = 2, MLIS_SUBMODEL_TAE = 3,
	MLIS_SUBMODEL_CLIP = 4, MLIS_SUBMODEL_CLIP2 = 5,
} MLIS_SubModel;

/* Option ids and their arguments are those of the reference (mlimgsynth.h:174-346). */
typedef enum MLIS_Option {
	MLIS_OPT_NONE = 0, MLIS_OPT_BACKEND = 1, MLIS_OPT_MODEL = 2, MLIS_OPT_TAE = 3, MLIS_OPT_LORA_DIR = 4,
	MLIS_OPT_LORA = 5, MLIS_OPT_LORA_CLEAR = 6, MLIS_OPT_PROMPT = 7, MLIS_OPT_NPROMPT = 8, MLIS_OPT_IMAGE_DIM = 9,
	MLIS_OPT_BATCH_SIZE = 10, MLIS_OPT_CLIP_SKIP = 11, MLIS_OPT_CFG_SCALE = 12, MLIS_OPT_METHOD = 13,
	MLIS_OPT_SCHEDULER = 14, MLIS_OPT_STEPS = 15, MLIS_OPT_F_T_INI = 16, MLIS_OPT_F_T_END = 17,
	MLIS_OPT_S_NOISE = 18, MLIS_OPT_S_ANCESTRAL = 19, MLIS_OPT_IMAGE = 20, MLIS_OPT_IMAGE_MASK = 21,
	MLIS_OPT_NO_DECODE = 22, MLIS_OPT_TENSOR_USE_FLAGS = 23, MLIS_OPT_SEED = 24, MLIS_OPT_VAE_TILE = 25,
	MLIS_OPT_UNET_SPLIT = 26, MLIS_OPT_THREADS = 27, MLIS_OPT_DUMP_FLAGS = 28, MLIS_OPT_AUX_DIR = 29,
	MLIS_OPT_CALLBACK = 30, MLIS_OPT_ERROR_HANDLER = 31, MLIS_OPT_LOG_LEVEL = 32, MLIS_OPT_MODEL_TYPE = 33,
	MLIS_OPT_WEIGHT_TYPE = 34, MLIS_OPT_NO_PROMPT_PARSE = 35, MLIS_OPT__LAST = 35,
} MLIS_Option;

typedef struct MLIS_Ctx MLIS_Ctx;

typedef struct MLIS_Image { uint8_t* d; size_t sz; unsigned w, h, c; int flags; } MLIS_Image;

typedef struct MLIS_Progress { MLIS_Stage stage; int step, step_end, nfe; double step_time; double time; } MLIS_Progress;

typedef struct MLIS_ErrorInfo { MLIS_ErrorCode code; const char* desc; } MLIS_ErrorInfo;

typedef struct MLIS_BackendInfo {
	const char* name; unsigned n_dev;
	struct MLIS_BackendDeviceInfo { const char *name, *desc; size_t mem_free, mem_total; } *devs;
} MLIS_BackendInfo;

typedef struct MLIS_Tensor { float* d; int n[4]; int flags; } MLIS_Tensor;

typedef int (*MLIS_Callback)(void*, MLIS_Ctx*, const MLIS_Progress*);
typedef void (*MLIS_ErrorHandler)(void*, MLIS_Ctx*, const MLIS_ErrorInfo*);

#define mlis_ctx_create()  mlis_ctx_create_i(MLIS_VERSION)
MLIS_Ctx* mlis_ctx_create_i(int version);
void mlis_ctx_destroy(MLIS_Ctx** pctx);
const char* mlis_errstr_get(const MLIS_Ctx* ctx);

int mlis_option_set(MLIS_Ctx* ctx, MLIS_Option id, ...);
int mlis_option_set_str(MLIS_Ctx* ctx, const char* name, const char* value);
int mlis_option_get(MLIS_Ctx* ctx, MLIS_Option id, ...);

int mlis_setup(MLIS_Ctx* ctx);
int mlis_generate(MLIS_Ctx* ctx);
MLIS_Image* mlis_image_get(MLIS_Ctx* ctx, int idx);
const char* mlis_infotext_get(MLIS_Ctx* ctx, int idx);
MLIS_Tensor* mlis_tensor_get(MLIS_Ctx* ctx, MLIS_TensorId id);
const MLIS_BackendInfo* mlis_backend_info_get(MLIS_Ctx* ctx, unsigned idx, int flags);

const char* mlis_stage_str(MLIS_Stage id);
const char* mlis_stage_desc(MLIS_Stage id);
MLIS_Stage mlis_stage_fromz(const char* str);
const char* mlis_loglvl_str(MLIS_LogLvl id);
MLIS_LogLvl mlis_loglvl_fromz(const char* str);
const char* mlis_method_str(MLIS_Method id);
MLIS_Method mlis_method_fromz(const char* str);
const char* mlis_sched_str(MLIS_Scheduler id);
MLIS_Scheduler mlis_sched_fromz(const char* str);
const char* mlis_model_type_str(MLIS_ModelType id);
const char* mlis_model_type_desc(MLIS_ModelType id);
MLIS_ModelType mlis_model_type_fromz(const char* str);
const char* mlis_option_str(MLIS_Option id);
MLIS_Option mlis_option_fromz(const char* str);

int mlis_image_encode(MLIS_Ctx* ctx, const MLIS_Tensor* image, MLIS_Tensor* latent, int flags);
int mlis_image_decode(MLIS_Ctx* ctx, const MLIS_Tensor* latent, MLIS_Tensor* image, int flags);
int mlis_mask_encode(MLIS_Ctx* ctx, const MLIS_Tensor* mask, MLIS_Tensor* lmask, int flags);
int mlis_text_tokenize(MLIS_Ctx* ctx, const char* text, int32_t** ptokens, MLIS_SubModel model);
int mlis_clip_text_encode(MLIS_Ctx* ctx, const char* text, MLIS_Tensor* embed, MLIS_Tensor* feat, MLIS_SubModel model, int flags);
enum { MLIS_CTEF_NO_NORM = 1 };

void   mlis_tensor_free(MLIS_Tensor*);
size_t mlis_tensor_count(const MLIS_Tensor*);
void   mlis_tensor_resize(MLIS_Tensor*, int n0, int n1, int n2, int n3);
void   mlis_tensor_resize_like(MLIS_Tensor*, const MLIS_Tensor*);
void   mlis_tensor_copy(MLIS_Tensor*, const MLIS_Tensor*);
float  mlis_tensor_similarity(const MLIS_Tensor*, const MLIS_Tensor*);

/* ---- additions (not in the reference header) ---- */
/* One UNet evaluation through the library's denoiser on caller data: x [w,h,4,n_img], cond
 * [n_ctx,77,n_img], label [adm,n_img] or NULL, sigma -> dx (eps / v converted like unet.c:460-497).
 * Used by the parity tests and the benchmark's per-NFE timing. */
int mlis_unet_eval(MLIS_Ctx* ctx, const MLIS_Tensor* x, const MLIS_Tensor* cond, const MLIS_Tensor* label, float sigma, MLIS_Tensor* dx);

/* VAE decode tiles spread across GPUs (one process per GPU). The reference decodes the tiles of `vae_tile` serially
 * (vae.c:331-391); they are independent, so rank r of `world` decodes tiles r, r + world, ... of the reference's row-major
 * tile list into consecutive slots of a CALLER-owned device buffer (f32, tile_w_px * tile_h_px * 3 floats per slot,
 * ceil(n_tiles / world) slots). The caller gathers the buffers of all ranks on one rank (NCCL, device to device; layout
 * [world][slots_per_worker][slot]) and calls merge there: tiles are pasted in the reference's order (later tiles
 * overwrite earlier ones, vae.c:365-387), then (x+1)/2, RGB8 pack and the usual mlis_image_get()/MLIS_TENSOR_IMAGE.
 * With world = 1 this is the serial tiled decode. */
int mlis_b200_vae_tile_plan(MLIS_Ctx* ctx, int lw, int lh, int* n_tiles, int* tile_w_px, int* tile_h_px);
int mlis_b200_vae_tiles_decode(MLIS_Ctx* ctx, const MLIS_Tensor* latent, int rank, int world, float* tiles_dev);
int mlis_b200_vae_tiles_merge(MLIS_Ctx* ctx, int lw, int lh, const float* gathered_dev, int world, int slots_per_worker, MLIS_Tensor* image);

/* Cross-GPU CFG split (opt-in, for fewer images than GPUs; the default batches both halves on one GPU). The reference
 * evaluates the conditional and the unconditional UNet one after the other (mlimgsynth.c:1578-1583); they only share x and
 * sigma, so a PAIR of contexts on two GPUs -- same model, options, seed and prompts -- can each evaluate one half
 * (half 0 = conditional, 1 = unconditional). Once per UNet evaluation the library calls `exchange` with this rank's output
 * (device pointer, n floats) and a device buffer to fill with the peer's output (a 2-rank all-gather, e.g. NCCL); both ranks
 * then form the same dx and keep identical sampler states. half < 0 switches the mode off. */
typedef int (*MLIS_B200_CfgExchange)(void* user, const float* mine_dev, float* other_dev, size_t n);
int mlis_b200_cfg_split_set(MLIS_Ctx* ctx, int half, MLIS_B200_CfgExchange exchange, void* user);

/* Device pointer of the RGB8 images ([n][h][w][3] bytes) of the last generation / decode, for device-to-device gathers. */
int mlis_b200_images_device(MLIS_Ctx* ctx, const uint8_t** dev, int* w, int* h, int* n);

#ifdef __cplusplus
}
#endif
#endif
