/* mlimgsynth.c -- the public mlis_* API of the B200 host layer (see include/mlimgsynth_b200.h).
 *
 * Orchestration follows the reference's mlimgsynth.c (mlis_setup :1251, cond encode :1502-1563,
 * mlis_denoise_dxdt :1572, mlis_generate :1634) with a device-first data flow:
 *   - the latent batch lives in HBM from the initial noise to the RGB8 pack; per UNet evaluation
 *     the host issues two fused element-wise launches and one CUDA-graph replay;
 *   - cond / uncond halves of classifier-free guidance run as ONE batched UNet evaluation;
 *   - CLIP / UNet / VAE graphs and their weights stay resident between generations.
 */
#include "mlimgsynth_b200.h"
#include "mlblock.h"
#include "unet.h"
#include "vae.h"
#include "clip.h"
#include "sampling.h"
#include "prompt_preproc.h"
#include "lora.h"
#include <ctype.h>
#include <math.h>
#include <time.h>

#define CTX_SIGNATURE 0x4D4C4942u   /* "MLIB" */
enum { CF_NO_DECODE = 1, CF_USE_TAE = 2, CF_NO_PROMPT_PARSE = 4, CF_MODEL_TYPE_SET = 8, CF_WEIGHT_TYPE_SET = 16 };
enum { RDY_BACKEND = 1, RDY_MODEL = 2, RDY_LORAS = 4 };

typedef struct LoraCfg { char* path; float mult; int from_prompt; } LoraCfg;

struct MLIS_Ctx {
	uint32_t signature;
	/* configuration */
	char *backend_name, *path_model, *path_tae, *lora_dir, *aux_dir, *prompt_raw, *nprompt_raw;
	LoraCfg* loras; int n_loras, cap_loras;
	PromptText prompt, nprompt;
	int width, height, n_batch, clip_skip, vae_tile, flags, tuflags, model_type, rflags;
	float cfg_scale;
	MLIS_Callback callback; void* callback_user;
	MLIS_ErrorHandler errh; void* errh_user;
	int cfg_half; MLIS_B200_CfgExchange cfg_exchange; void* cfg_exchange_user;   /* cross-GPU CFG split (opt-in), -1 = off */
	float* cfg_other_dev; size_t cfg_other_n;
	/* runtime */
	ggml_backend_t backend;
	TStore tstore; bool tstore_open;
	enum ggml_type wtype;
	const UnetParams* unet_p; const VaeParams* vae_p; const SdTaeParams* tae_p; const ClipParams *clip_p, *clip2_p;
	MLCtx ctx_clip, ctx_clip2, ctx_clip2f, ctx_unet, ctx_vdec, ctx_venc, ctx_tdec, ctx_tenc;
	ClipState st_clip, st_clip2, st_clip2f;
	UnetState unet;
	CodecState st_vdec, st_venc, st_tdec, st_tenc;
	DenoiseSampler sampler;
	RngPhilox rngs[64];
	/* tensors */
	HTensor image, mask, latent, lmask, cond, label, ncond, nlabel, tmp[8];
	float *latent_dev, *image_dev, *lmask_dev; uint8_t* u8_dev; size_t latent_dev_n, image_dev_n, lmask_dev_n, u8_dev_n;
	bool latent_host_stale, image_host_stale;
	MLIS_Image imgex; uint8_t* imgex_all; size_t imgex_cap; int img_w, img_h, img_n;      /* imgex_all: page-locked */
	MLIS_Progress prg; double t_last;
	int32_t* tokens; int n_tokens, cap_tokens; float* tok_w; int cap_tok_w;
	char* infotext;
	char errstr[512];
	MLIS_BackendInfo backend_info; struct MLIS_BackendDeviceInfo devinfo[16];
};

/* The reference keeps ONE process-wide noise stream (rng_philox.c:7 g_rng; seed defaults to the
 * wall clock at first context creation, setting the seed does not reset the offset). */
static RngPhilox g_rng;
static bool g_rng_seeded;

/* ------------------------------------------------------------------ enum <-> string */
static const char* k_methods[] = { "none", "euler", "heun", "taylor3", "dpmpp2m", "dpmpp2s" };
static const char* k_scheds[] = { "none", "uniform", "karras" };
static const char* k_models[] = { "none", "sd1", "sd2", "sdxl" };
static const char* k_stages[] = { "idle", "cond_encode", "image_encode", "image_decode", "denoise" };
static const char* k_options[] = { "none", "backend", "model", "tae", "lora_dir", "lora", "lora_clear", "prompt", "nprompt",
	"image_dim", "batch_size", "clip_skip", "cfg_scale", "method", "scheduler", "steps", "f_t_ini", "f_t_end", "s_noise",
	"s_ancestral", "image", "image_mask", "no_decode", "tensor_use_flags", "seed", "vae_tile", "unet_split", "threads",
	"dump_flags", "aux_dir", "callback", "error_handler", "log_level", "model_type", "weight_type", "no_prompt_parse" };

/* identifier comparison as the reference header documents it (mlimgsynth.h:435-441: "CFG_SCALE" = "cfg_scale" =
 * "cfg-scale"): case-insensitive, '-' == '_', '+' == 'p' (for "dpm++2m"). The reference's strsl_cmpz_id
 * (mlimgsynth.c:157-170) intends the same; its upper-case branch subtracts 'A' without adding 'a' back, so there only
 * lower-case names match -- every name the reference accepts is accepted here with the same meaning. */
static bool id_eq(const char* s, const char* name)
{
	for (; *s && *name; ++s, ++name) {
		char c = *s == '-' ? '_' : *s == '+' ? 'p' : (*s >= 'A' && *s <= 'Z') ? (char)(*s - 'A' + 'a') : *s;
		if (c != *name) return false;
	}
	return !*s && !*name;
}
static int enum_from(const char* s, const char* const* names, int n, int dflt)
{
	for (int i = 0; i < n; ++i) if (id_eq(s, names[i])) return i;
	return dflt;
}
#define NAMES(a) a, (int)(sizeof(a) / sizeof(a[0]))
static const char* k_stage_desc[] = { "Idle", "Conditioning encoding", "Image encoding", "Image decoding", "Denoising" };
static const char* k_model_desc[] = { "None", "Stable Diffusion 1.x", "Stable Diffusion 2.x", "Stable Diffusion XL" };
static const struct { const char* name; int id; } k_loglvls[] = {
	{ "none", MLIS_LOGLVL_NONE }, { "error", MLIS_LOGLVL_ERROR }, { "warning", MLIS_LOGLVL_WARNING }, { "info", MLIS_LOGLVL_INFO },
	{ "verbose", MLIS_LOGLVL_VERBOSE }, { "debug", MLIS_LOGLVL_DEBUG }, { "max", MLIS_LOGLVL_MAX } };
const char* mlis_stage_str(MLIS_Stage id) { return (unsigned)id < 5 ? k_stages[id] : "???"; }
const char* mlis_stage_desc(MLIS_Stage id) { return (unsigned)id < 5 ? k_stage_desc[id] : "???"; }
MLIS_Stage mlis_stage_fromz(const char* s) { return (MLIS_Stage)enum_from(s, NAMES(k_stages), -1); }
const char* mlis_model_type_desc(MLIS_ModelType id) { return (unsigned)id < 4 ? k_model_desc[id] : "???"; }
const char* mlis_loglvl_str(MLIS_LogLvl id)
{
	for (unsigned i = 0; i < sizeof(k_loglvls) / sizeof(k_loglvls[0]); ++i) if ((int)id == k_loglvls[i].id) return k_loglvls[i].name;
	return "???";
}
MLIS_LogLvl mlis_loglvl_fromz(const char* s)
{
	for (unsigned i = 0; i < sizeof(k_loglvls) / sizeof(k_loglvls[0]); ++i) if (id_eq(s, k_loglvls[i].name)) return (MLIS_LogLvl)k_loglvls[i].id;
	return (MLIS_LogLvl)-1;
}
const char* mlis_method_str(MLIS_Method id) { return (unsigned)id < 6 ? k_methods[id] : "???"; }
MLIS_Method mlis_method_fromz(const char* s) { return (MLIS_Method)enum_from(s, NAMES(k_methods), -1); }
const char* mlis_sched_str(MLIS_Scheduler id) { return (unsigned)id < 3 ? k_scheds[id] : "???"; }
MLIS_Scheduler mlis_sched_fromz(const char* s) { return (MLIS_Scheduler)enum_from(s, NAMES(k_scheds), -1); }
const char* mlis_model_type_str(MLIS_ModelType id) { return (unsigned)id < 4 ? k_models[id] : "???"; }
MLIS_ModelType mlis_model_type_fromz(const char* s) { return (MLIS_ModelType)enum_from(s, NAMES(k_models), -1); }
const char* mlis_option_str(MLIS_Option id) { return (unsigned)id <= MLIS_OPT__LAST ? k_options[id] : "???"; }
MLIS_Option mlis_option_fromz(const char* s) { return (MLIS_Option)enum_from(s, NAMES(k_options), -1); }   /* -1 when unknown, as mlimgsynth.c:301 */

/* ------------------------------------------------------------------ errors, callbacks */
static int fail(MLIS_Ctx* S, int code, const char* where)
{
	snprintf(S->errstr, sizeof(S->errstr), "%s: %s", where, mlis_err_get());
	if (S->errh) { MLIS_ErrorInfo ei = { code, S->errstr }; S->errh(S->errh_user, S, &ei); }
	return code;
}
#define API_TRY(expr, where) do { int r_ = (expr); if (r_ < 0) return fail(S, r_, where); } while (0)

const char* mlis_errstr_get(const MLIS_Ctx* S) { return S->errstr; }

static int progress(MLIS_Ctx* S, MLIS_Stage stage, int step, int step_end)
{
	double t = time_now();
	S->prg.stage = stage; S->prg.step = step; S->prg.step_end = step_end;
	S->prg.step_time = t - S->t_last; S->prg.time = t; S->t_last = t;
	if (S->callback) { int r = S->callback(S->callback_user, S, &S->prg); if (r < 0) { mlis_err_set("aborted by the progress callback"); return r; } }
	return 1;
}

/* ------------------------------------------------------------------ context */
static void set_str(char** dst, const char* s) { free(*dst); *dst = s ? xstrdup(s) : NULL; }

MLIS_Ctx* mlis_ctx_create_i(int version)
{
	(void)version;
	MLIS_Ctx* S = xcalloc(1, sizeof(*S));
	S->signature = CTX_SIGNATURE;
	unet_params_init();            /* sigma tables are needed before the first UNet graph exists (dnsamp_init) */
	S->cfg_scale = 7;              /* mlimgsynth.c:474 */
	S->cfg_half = -1;
	S->n_batch = 1;
	S->wtype = GGML_TYPE_F16;
	if (!g_rng_seeded) {
		struct timespec ts; clock_gettime(CLOCK_REALTIME, &ts);
		g_rng.seed = (uint64_t)ts.tv_sec * 1000 + ts.tv_nsec / 1000000; g_rng.offset = 0; g_rng_seeded = true;
	}
	return S;
}

static void graphs_free(MLIS_Ctx* S)
{
	MLCtx* all[] = { &S->ctx_clip, &S->ctx_clip2, &S->ctx_clip2f, &S->ctx_unet, &S->ctx_vdec, &S->ctx_venc, &S->ctx_tdec, &S->ctx_tenc };
	for (unsigned i = 0; i < sizeof(all) / sizeof(all[0]); ++i) mlctx_end(all[i]);
	memset(&S->st_clip, 0, sizeof(S->st_clip)); memset(&S->st_clip2, 0, sizeof(S->st_clip2)); memset(&S->st_clip2f, 0, sizeof(S->st_clip2f));
	memset(&S->unet, 0, sizeof(S->unet));
	memset(&S->st_vdec, 0, sizeof(S->st_vdec)); memset(&S->st_venc, 0, sizeof(S->st_venc));
	memset(&S->st_tdec, 0, sizeof(S->st_tdec)); memset(&S->st_tenc, 0, sizeof(S->st_tenc));
}

static void loras_free(MLIS_Ctx* S, bool only_prompt)
{
	int k = 0;
	for (int i = 0; i < S->n_loras; ++i) {
		if (only_prompt && !S->loras[i].from_prompt) { S->loras[k++] = S->loras[i]; continue; }
		free(S->loras[i].path);
	}
	if (k != S->n_loras) S->rflags &= ~RDY_LORAS;
	S->n_loras = k;
}

void mlis_ctx_destroy(MLIS_Ctx** pS)
{
	MLIS_Ctx* S = pS ? *pS : NULL;
	if (!S) return;
	if (S->backend) {
		graphs_free(S);
		dnsamp_free(&S->sampler);
		ggml_b200_free(S->latent_dev); ggml_b200_free(S->image_dev); ggml_b200_free(S->lmask_dev); ggml_b200_free(S->u8_dev); ggml_b200_free(S->cfg_other_dev);
		ggml_backend_free(S->backend);
	}
	if (S->tstore_open) tstore_free(&S->tstore);
	loras_free(S, false); free(S->loras);
	prompt_text_free(&S->prompt); prompt_text_free(&S->nprompt);
	HTensor* ts[] = { &S->image, &S->mask, &S->latent, &S->lmask, &S->cond, &S->label, &S->ncond, &S->nlabel };
	for (unsigned i = 0; i < 8; ++i) ht_free(ts[i]);
	for (int i = 0; i < 8; ++i) ht_free(&S->tmp[i]);
	free(S->backend_name); free(S->path_model); free(S->path_tae); free(S->lora_dir); free(S->aux_dir);
	free(S->prompt_raw); free(S->nprompt_raw); free(S->tokens); free(S->tok_w); free(S->infotext); ggml_b200_host_free(S->imgex_all);
	free(S);
	*pS = NULL;
}

/* ------------------------------------------------------------------ options */
static int model_type_set(MLIS_Ctx* S, int mt)
{
	S->model_type = mt;
	S->vae_p = &g_vae_sd1; S->tae_p = &g_sdtae_sd1; S->clip2_p = NULL;
	switch (mt) {
	case MLIS_MODEL_TYPE_SD1: S->unet_p = &g_unet_sd1; S->clip_p = &g_clip_vit_l_14; if (!S->clip_skip) S->clip_skip = 1; break;
	case MLIS_MODEL_TYPE_SD2: S->unet_p = &g_unet_sd2; S->clip_p = &g_clip_vit_h_14; if (!S->clip_skip) S->clip_skip = 2; break;
	case MLIS_MODEL_TYPE_SDXL: S->unet_p = &g_unet_sdxl; S->vae_p = &g_vae_sdxl; S->clip_p = &g_clip_vit_l_14; S->clip2_p = &g_clip_vit_bigg_14;
		if (!S->clip_skip) S->clip_skip = 2;
		break;
	default: FAIL(MLIS_E_OPT_VALUE, "invalid model type %d", mt);
	}
	if (!S->width) { int d = mt == MLIS_MODEL_TYPE_SD1 ? 512 : mt == MLIS_MODEL_TYPE_SD2 ? 768 : 1024; S->width = S->height = d; }
	return 1;
}

static int lora_add(MLIS_Ctx* S, const char* name, size_t len, float mult, int from_prompt)
{
	if (!(mult >= 0 && mult <= 1)) FAIL(MLIS_E_OPT_VALUE, "lora multiplier out of range [0,1]");
	/* the option accepts 0 (options_set.c.h:38) but the merge requires scale > 0 (lora.c:42): a zero-weight LoRA is a no-op */
	if (mult == 0) { log_warn("LoRA '%.*s' has multiplier 0: ignored", (int)len, name); return 1; }
	char path[1024];
	bool is_path = memchr(name, '/', len) != NULL || (len > 12 && !memcmp(name + len - 12, ".safetensors", 12));
	if (is_path || !S->lora_dir) snprintf(path, sizeof(path), "%.*s", (int)len, name);
	else snprintf(path, sizeof(path), "%s/%.*s.safetensors", S->lora_dir, (int)len, name);
	LoraCfg l = { xstrdup(path), mult, from_prompt };
	ARR_PUSH(S->loras, S->n_loras, S->cap_loras, l);
	S->rflags &= ~RDY_LORAS;
	return 1;
}

static int prompt_set(MLIS_Ctx* S, PromptText* pt, char** raw, const char* text)
{
	set_str(raw, text);
	if (S->flags & CF_NO_PROMPT_PARSE) { prompt_text_set_raw(pt, text, strlen(text)); return 1; }
	CHECK(prompt_text_set_parse(pt, text, strlen(text)));
	for (int i = 0; i < pt->n_loras; ++i) CHECK(lora_add(S, pt->data + pt->loras[i].beg, pt->loras[i].len, pt->loras[i].w, 1));
	return 1;
}

static void image_to_tensors(MLIS_Ctx* S, const MLIS_Image* img, bool mask_only)
{
	/* u8 -> f32 * (1/255), planar (mlimgsynth.c:131-151); alpha channel (c == 4 or 2) becomes the mask */
	int w = img->w, h = img->h, c = img->c;
	if (mask_only) {
		ht_resize(&S->mask, w, h, 1, 1);
		for (int i = 0; i < w * h; ++i) S->mask.d[i] = img->d[(size_t)i * c] * (1 / 255.0f);
		S->tuflags |= MLIS_TUF_MASK;
		return;
	}
	int cc = c >= 3 ? 3 : 1;
	ht_resize(&S->image, w, h, 3, 1);
	for (int k = 0; k < 3; ++k)
		for (int i = 0; i < w * h; ++i) S->image.d[(size_t)k * w * h + i] = img->d[(size_t)i * c + (cc == 3 ? k : 0)] * (1 / 255.0f);
	S->tuflags |= MLIS_TUF_IMAGE;
	S->width = w; S->height = h;
	if (c == 4 || c == 2) {
		ht_resize(&S->mask, w, h, 1, 1);
		for (int i = 0; i < w * h; ++i) S->mask.d[i] = img->d[(size_t)i * c + c - 1] * (1 / 255.0f);
		S->tuflags |= MLIS_TUF_MASK;
	}
}

/* typed value of one option argument */
typedef struct Arg { const char* s; long long i; double f; const void* p; void* p2; } Arg;

static int option_apply(MLIS_Ctx* S, MLIS_Option id, const Arg* a, int n_arg)
{
	switch (id) {
	case MLIS_OPT_BACKEND: set_str(&S->backend_name, a[0].s); S->rflags &= ~RDY_BACKEND; break;
	case MLIS_OPT_MODEL: set_str(&S->path_model, a[0].s); S->rflags &= ~(RDY_MODEL | RDY_LORAS); break;
	case MLIS_OPT_TAE: set_str(&S->path_tae, a[0].s && a[0].s[0] ? a[0].s : NULL);
		if (S->path_tae) S->flags |= CF_USE_TAE; else S->flags &= ~CF_USE_TAE;
		S->rflags &= ~(RDY_MODEL | RDY_LORAS); break;
	case MLIS_OPT_MODEL_TYPE: CHECK(model_type_set(S, (int)a[0].i)); S->flags |= CF_MODEL_TYPE_SET; break;
	case MLIS_OPT_AUX_DIR: set_str(&S->aux_dir, a[0].s); break;
	case MLIS_OPT_LORA_DIR: set_str(&S->lora_dir, a[0].s); break;
	case MLIS_OPT_LORA: CHECK(lora_add(S, a[0].s, strlen(a[0].s), n_arg > 1 ? (float)a[1].f : 1.0f, 0)); break;
	case MLIS_OPT_LORA_CLEAR: loras_free(S, false); S->rflags &= ~RDY_LORAS; break;
	case MLIS_OPT_PROMPT: CHECK(prompt_set(S, &S->prompt, &S->prompt_raw, a[0].s)); break;
	case MLIS_OPT_NPROMPT: CHECK(prompt_set(S, &S->nprompt, &S->nprompt_raw, a[0].s)); break;
	case MLIS_OPT_NO_PROMPT_PARSE: if (a[0].i) S->flags |= CF_NO_PROMPT_PARSE; else S->flags &= ~CF_NO_PROMPT_PARSE; break;
	case MLIS_OPT_IMAGE_DIM:
		if (a[0].i % 64 || a[1].i % 64) FAIL(MLIS_E_OPT_VALUE, "image dimensions must be multiples of 64");
		S->width = (int)a[0].i; S->height = (int)a[1].i; break;
	case MLIS_OPT_BATCH_SIZE: if (a[0].i < 0 || a[0].i > 32) FAIL(MLIS_E_OPT_VALUE, "batch size out of range (1-32)"); S->n_batch = a[0].i ? (int)a[0].i : 1; break;
	case MLIS_OPT_CLIP_SKIP: S->clip_skip = (int)a[0].i; break;
	case MLIS_OPT_CFG_SCALE: S->cfg_scale = (float)a[0].f; break;
	case MLIS_OPT_METHOD: if (a[0].i < 1 || a[0].i > 5) FAIL(MLIS_E_OPT_VALUE, "invalid method"); S->sampler.c.method = (int)a[0].i; break;
	case MLIS_OPT_SCHEDULER: if (a[0].i < 1 || a[0].i > 2) FAIL(MLIS_E_OPT_VALUE, "invalid scheduler"); S->sampler.c.sched = (int)a[0].i; break;
	case MLIS_OPT_STEPS: S->sampler.c.n_step = (int)a[0].i; break;
	case MLIS_OPT_F_T_INI: S->sampler.c.f_t_ini = (float)a[0].f; break;
	case MLIS_OPT_F_T_END: S->sampler.c.f_t_end = (float)a[0].f; break;
	case MLIS_OPT_S_NOISE: S->sampler.c.s_noise = (float)a[0].f; break;
	case MLIS_OPT_S_ANCESTRAL: S->sampler.c.s_ancestral = (float)a[0].f; break;
	case MLIS_OPT_IMAGE: image_to_tensors(S, a[0].p, false); break;
	case MLIS_OPT_IMAGE_MASK: image_to_tensors(S, a[0].p, true); break;
	case MLIS_OPT_NO_DECODE: if (a[0].i) S->flags |= CF_NO_DECODE; else S->flags &= ~CF_NO_DECODE; break;
	case MLIS_OPT_TENSOR_USE_FLAGS: S->tuflags = (int)a[0].i; break;
	case MLIS_OPT_SEED: g_rng.seed = (uint64_t)a[0].i;   /* the offset keeps counting (options_set.c.h:162-168) ... */
		if (n_arg > 1) g_rng.offset = (uint32_t)a[1].i;       /* ... unless given explicitly: "seed", "42,0" (addition, string form only) */
		break;
	case MLIS_OPT_VAE_TILE: S->vae_tile = (int)a[0].i; break;
	case MLIS_OPT_UNET_SPLIT: case MLIS_OPT_THREADS: case MLIS_OPT_DUMP_FLAGS: break;   /* accepted, no effect on B200 */
	case MLIS_OPT_WEIGHT_TYPE:
		if (a[0].i != GGML_TYPE_F16 && a[0].i != GGML_TYPE_F32) FAIL(MLIS_E_OPT_VALUE, "weight type must be f16 or f32 on the B200 engine");
		S->wtype = (enum ggml_type)a[0].i; S->flags |= CF_WEIGHT_TYPE_SET; break;
	case MLIS_OPT_CALLBACK: S->callback = (MLIS_Callback)a[0].p; S->callback_user = a[1].p2; break;
	case MLIS_OPT_ERROR_HANDLER: S->errh = (MLIS_ErrorHandler)a[0].p; S->errh_user = a[1].p2; break;
	case MLIS_OPT_LOG_LEVEL: {
		int l = (int)a[0].i;
		/* increase starts directly from INFO (options_set.c.h:229-238) */
		if ((l & 0xf00) == 0x100) { if (g_log_level < LOG_INFO) g_log_level = LOG_INFO; else g_log_level += l & 0xff; }
		else if ((l & 0xf00) == 0x200) g_log_level -= l & 0xff; else g_log_level = l;
	} break;
	default: FAIL(MLIS_E_UNK_OPT, "unknown option %d", (int)id);
	}
	return 1;
}

/* argument kinds per option: s string, i int, f double, u uint64, p pointer, e enum (int / name) */
static const char* option_sig(MLIS_Option id)
{
	switch (id) {
	case MLIS_OPT_BACKEND: return "sS";
	case MLIS_OPT_MODEL: case MLIS_OPT_TAE: case MLIS_OPT_AUX_DIR: case MLIS_OPT_LORA_DIR: case MLIS_OPT_PROMPT: case MLIS_OPT_NPROMPT: return "s";
	case MLIS_OPT_LORA: return "sF";
	case MLIS_OPT_LORA_CLEAR: return "";
	case MLIS_OPT_IMAGE_DIM: return "ii";
	case MLIS_OPT_CFG_SCALE: case MLIS_OPT_F_T_INI: case MLIS_OPT_F_T_END: case MLIS_OPT_S_NOISE: case MLIS_OPT_S_ANCESTRAL: return "f";
	case MLIS_OPT_SEED: return "uI";
	case MLIS_OPT_IMAGE: case MLIS_OPT_IMAGE_MASK: return "p";
	case MLIS_OPT_CALLBACK: case MLIS_OPT_ERROR_HANDLER: return "pp";
	case MLIS_OPT_METHOD: case MLIS_OPT_SCHEDULER: case MLIS_OPT_MODEL_TYPE: case MLIS_OPT_WEIGHT_TYPE: case MLIS_OPT_LOG_LEVEL: return "e";
	default: return "i";
	}
}

int mlis_option_set(MLIS_Ctx* S, MLIS_Option id, ...)
{
	if ((int)id <= 0 || (int)id > MLIS_OPT__LAST) { mlis_err_set("unknown option %d", (int)id); return fail(S, MLIS_E_UNK_OPT, "mlis_option_set"); }
	const char* sig = option_sig(id);
	Arg a[4]; memset(a, 0, sizeof(a));
	va_list ap; va_start(ap, id);
	int n = 0;
	for (; sig[n]; ++n) {
		switch (sig[n]) {
		case 's': case 'S': a[n].s = va_arg(ap, const char*); break;
		case 'i': case 'e': a[n].i = va_arg(ap, int); break;
		case 'I': goto done;   /* string-form extension only */
		case 'f': case 'F': a[n].f = va_arg(ap, double); break;
		case 'u': a[n].i = (long long)va_arg(ap, uint64_t); break;
		case 'p': if (n == 0) a[n].p = va_arg(ap, const void*); else a[n].p2 = va_arg(ap, void*); break;
		}
	}
done:
	va_end(ap);
	API_TRY(option_apply(S, id, a, n), "mlis_option_set");
	return 1;
}

int mlis_option_set_str(MLIS_Ctx* S, const char* name, const char* value)
{
	MLIS_Option id = mlis_option_fromz(name);
	if ((int)id <= 0) { mlis_err_set("unknown option '%s'", name); return fail(S, MLIS_E_UNK_OPT, "mlis_option_set_str"); }
	const char* sig = option_sig(id);
	Arg a[4]; memset(a, 0, sizeof(a));
	char buf[4][1024];
	int n = 0;
	const char* v = value ? value : "";
	if (id == MLIS_OPT_SEED && !*v) return 1;      /* empty string: keep the (random) seed, options_set.c.h:163-164 */
	bool whole = (id == MLIS_OPT_MODEL || id == MLIS_OPT_TAE || id == MLIS_OPT_AUX_DIR || id == MLIS_OPT_LORA_DIR ||
		id == MLIS_OPT_PROMPT || id == MLIS_OPT_NPROMPT);
	for (; sig[n]; ++n) {
		/* arguments are separated by ',' (mlimgsynth.c:844-863); path/prompt options take the whole value */
		const char* e = whole ? v + strlen(v) : strchr(v, ',');
		if (!e) e = v + strlen(v);
		bool present = e > v || (n == 0);
		bool optional = sig[n] == 'S' || sig[n] == 'F' || sig[n] == 'I';
		if (!present && optional) break;
		if (whole) a[n].s = v;
		else { snprintf(buf[n], sizeof(buf[n]), "%.*s", (int)(e - v), v); a[n].s = buf[n]; }
		char* tail = NULL;
		switch (sig[n]) {
		case 'i': case 'I': a[n].i = strtoll(a[n].s, &tail, 0);
			if (tail == a[n].s) { if (!strcmp(a[n].s, "true")) a[n].i = 1; else if (!strcmp(a[n].s, "false")) a[n].i = 0; else goto bad; }
			break;
		case 'u': a[n].i = (long long)strtoull(a[n].s, &tail, 0); if (tail == a[n].s) goto bad; break;
		case 'f': case 'F': a[n].f = strtod(a[n].s, &tail); if (tail == a[n].s) goto bad; break;
		case 'e': {
			int x = -1;
			if (id == MLIS_OPT_METHOD) x = mlis_method_fromz(a[n].s);
			else if (id == MLIS_OPT_SCHEDULER) x = mlis_sched_fromz(a[n].s);
			else if (id == MLIS_OPT_MODEL_TYPE) x = mlis_model_type_fromz(a[n].s);
			else if (id == MLIS_OPT_WEIGHT_TYPE) x = id_eq(a[n].s, "f16") ? GGML_TYPE_F16 : id_eq(a[n].s, "f32") ? GGML_TYPE_F32 : -1;
			else if (id == MLIS_OPT_LOG_LEVEL) x = mlis_loglvl_fromz(a[n].s);      /* level names (options_set.c.h:220-226) */
			if (x < 0) { x = (int)strtol(a[n].s, &tail, 0); if (tail == a[n].s) goto bad; }
			a[n].i = x;
		} break;
		case 'p': mlis_err_set("option '%s' cannot be set from a string", name); return fail(S, MLIS_E_OPT_VALUE, "mlis_option_set_str");
		}
		if (!*e) { n++; break; }
		v = e + 1;
	}
	API_TRY(option_apply(S, id, a, n), "mlis_option_set_str");
	return 1;
bad:
	mlis_err_set("option '%s': invalid value '%s'", name, value ? value : "");
	return fail(S, MLIS_E_OPT_VALUE, "mlis_option_set_str");
}

int mlis_option_get(MLIS_Ctx* S, MLIS_Option id, ...)
{
	va_list ap; va_start(ap, id);
	int r = 1;
	switch (id) {
	case MLIS_OPT_MODEL_TYPE: *va_arg(ap, int*) = S->model_type; break;
	case MLIS_OPT_IMAGE_DIM: *va_arg(ap, int*) = S->width; *va_arg(ap, int*) = S->height; break;
	case MLIS_OPT_BATCH_SIZE: *va_arg(ap, int*) = S->n_batch; break;
	case MLIS_OPT_CFG_SCALE: *va_arg(ap, double*) = S->cfg_scale; break;
	case MLIS_OPT_STEPS: *va_arg(ap, int*) = S->sampler.n_step; break;
	case MLIS_OPT_SEED: *va_arg(ap, uint64_t*) = g_rng.seed; break;
	case MLIS_OPT_CLIP_SKIP: *va_arg(ap, int*) = S->clip_skip; break;
	case MLIS_OPT_METHOD: *va_arg(ap, int*) = S->sampler.c.method; break;
	case MLIS_OPT_SCHEDULER: *va_arg(ap, int*) = S->sampler.c.sched; break;
	case MLIS_OPT_WEIGHT_TYPE: *va_arg(ap, int*) = S->wtype; break;
	default: mlis_err_set("option %d cannot be read", (int)id); r = MLIS_E_UNK_OPT;
	}
	va_end(ap);
	return r < 0 ? fail(S, r, "mlis_option_get") : r;
}

/* ------------------------------------------------------------------ setup */
static MLCtx* graph_ctx_init(MLIS_Ctx* S, MLCtx* C)
{
	C->backend = S->backend; C->tstore = &S->tstore; C->c.wtype = S->wtype;
	C->c.flags = g_log_level >= LOG_INFO ? 0 : MLB_F_QUIET;
	return C;
}

static int setup(MLIS_Ctx* S)
{
	if (!(S->rflags & RDY_BACKEND)) {
		if (S->backend) { graphs_free(S); ggml_backend_free(S->backend); S->backend = NULL; }
		S->backend = (S->backend_name && *S->backend_name) ? ggml_backend_init_by_name(S->backend_name, NULL) : ggml_backend_init_best();
		if (!S->backend) FAIL(MLIS_E_UNKNOWN, "B200 backend init failed (no sm_100 device? there is no CPU fallback)");
		log_info("Backend: %s", ggml_backend_name(S->backend));
		S->rflags |= RDY_BACKEND;
	}
	if (!(S->rflags & RDY_MODEL)) {
		graphs_free(S);
		if (S->tstore_open) { tstore_free(&S->tstore); S->tstore_open = false; }
		if (!S->path_model) FAIL(MLIS_E_UNKNOWN, "No model file set");
		double t0 = time_now();
		CHECK(tstore_read_safetensors(&S->tstore, S->path_model, tnconv_sd, NULL));
		S->tstore_open = true;
		if (S->path_tae) CHECK(tstore_read_safetensors(&S->tstore, S->path_tae, NULL, "tae."));
		log_info("Model header loaded: %d tensors {%.3fs}", S->tstore.n, time_now() - t0);
		/* model type from the cross-attention key width (mlimgsynth.c:1207-1249) */
		int mt = 0; const TSEntry* te;
		if ((te = tstore_find(&S->tstore, "unet.in.1.1.transf.0.attn2.k_proj.weight"))) mt = te->shape[0] == 768 ? MLIS_MODEL_TYPE_SD1 : te->shape[0] == 1024 ? MLIS_MODEL_TYPE_SD2 : 0;
		else if ((te = tstore_find(&S->tstore, "unet.in.4.1.transf.0.attn2.k_proj.weight"))) mt = te->shape[0] == 2048 ? MLIS_MODEL_TYPE_SDXL : 0;
		if (mt) CHECK(model_type_set(S, mt));
		else if (!(S->flags & CF_MODEL_TYPE_SET)) FAIL(-1, "could not detect the model type");
		if (te && !(S->flags & CF_WEIGHT_TYPE_SET)) S->wtype = te->dtype == TS_F32 ? GGML_TYPE_F32 : GGML_TYPE_F16;
		log_info("Model type: %s", mlis_model_type_str(S->model_type));
		S->rflags |= RDY_MODEL;
		S->rflags &= ~RDY_LORAS;
	}
	if (!(S->rflags & RDY_LORAS)) {
		/* merges are cumulative on the store: start again from the file when the set changes */
		if (S->n_loras || S->tstore.n) {
			bool dirty = false;
			for (int i = 0; i < S->tstore.n; ++i) if (S->tstore.e[i].owned) dirty = true;
			if (dirty) { S->rflags &= ~RDY_MODEL; return setup(S); }
		}
		double t0 = time_now();
		for (int i = 0; i < S->n_loras; ++i) {
			TStore tl; memset(&tl, 0, sizeof(tl));
			int rr = tstore_read_safetensors(&tl, S->loras[i].path, lora_name_conv, NULL);
			if (rr < 0) { tstore_free(&tl); return rr; }
			int r = lora_apply(&S->tstore, &tl, S->loras[i].mult, S->wtype == GGML_TYPE_F32 ? TS_F32 : TS_F16);
			tstore_free(&tl);
			if (r < 0) return r;
			log_info("LoRA '%s' x%g: %d tensors merged", S->loras[i].path, S->loras[i].mult, r);
		}
		if (S->n_loras) { log_info("LoRA's applied: %d {%.3fs}", S->n_loras, time_now() - t0); graphs_free(S); }
		S->rflags |= RDY_LORAS;
	}
	return 1;
}

int mlis_setup(MLIS_Ctx* S)
{
	if (!S || S->signature != CTX_SIGNATURE) return -1;
	API_TRY(setup(S), "mlis_setup");
	return 1;
}

const MLIS_BackendInfo* mlis_backend_info_get(MLIS_Ctx* S, unsigned idx, int flags)
{
	(void)flags;
	if (idx >= ggml_backend_reg_count()) return NULL;
	ggml_backend_reg_t br = ggml_backend_reg_get(idx);
	MLIS_BackendInfo* bi = &S->backend_info;
	bi->name = ggml_backend_reg_name(br);
	bi->n_dev = (unsigned)ggml_backend_reg_dev_count(br);
	if (bi->n_dev > 16) bi->n_dev = 16;
	bi->devs = S->devinfo;
	for (unsigned i = 0; i < bi->n_dev; ++i) {
		ggml_backend_dev_t d = ggml_backend_reg_dev_get(br, i);
		S->devinfo[i].name = ggml_backend_dev_name(d);
		S->devinfo[i].desc = ggml_backend_dev_description(d);
		ggml_backend_dev_memory(d, &S->devinfo[i].mem_free, &S->devinfo[i].mem_total);
	}
	return bi;
}

/* ------------------------------------------------------------------ conditioning */
static int prompt_tokenize(MLIS_Ctx* S, const PromptText* p, const ClipParams* cp)
{
	CHECK(clip_tokenizer_load(S->aux_dir));
	S->n_tokens = 0;
	for (int i = 0; i < p->n_chunks; ++i) {
		int n0 = S->n_tokens;
		CHECK(clip_tokenize(cp, p->text + p->chunks[i].beg, p->chunks[i].len, &S->tokens, &S->n_tokens, &S->cap_tokens));
		if (S->n_tokens > S->cap_tok_w) { S->cap_tok_w = S->n_tokens * 2; S->tok_w = xrealloc(S->tok_w, S->cap_tok_w * sizeof(float)); }
		for (int k = n0; k < S->n_tokens; ++k) S->tok_w[k] = p->chunks[i].w;
	}
	log_info("Prompt: %d tokens", S->n_tokens);
	return S->n_tokens;
}

int mlis_text_tokenize(MLIS_Ctx* S, const char* text, int32_t** ptokens, MLIS_SubModel model)
{
	if (!S->clip_p) API_TRY(setup(S), "mlis_text_tokenize");
	const ClipParams* cp = model == MLIS_SUBMODEL_CLIP ? S->clip_p : model == MLIS_SUBMODEL_CLIP2 ? S->clip2_p : NULL;
	if (!cp) { mlis_err_set("invalid model for text tokenization: %d", (int)model); return fail(S, MLIS_E_UNKNOWN, "mlis_text_tokenize"); }
	prompt_text_set_raw(&S->prompt, text, strlen(text));
	int n = prompt_tokenize(S, &S->prompt, cp);
	if (n < 0) return fail(S, n, "mlis_text_tokenize");
	if (ptokens) *ptokens = S->tokens;
	return n;
}

static int clip_tokens_encode(MLIS_Ctx* S, int n_tok, const int32_t* toks, const float* weights, HTensor* embed, HTensor* feat,
	MLIS_SubModel model, int flags)
{
	CHECK(setup(S));
	bool second = model == MLIS_SUBMODEL_CLIP2;
	const ClipParams* cp = second ? S->clip2_p : model == MLIS_SUBMODEL_CLIP ? S->clip_p : NULL;
	if (!cp) FAIL(MLIS_E_UNKNOWN, "invalid model for text encoding: %d", (int)model);
	MLCtx* C = graph_ctx_init(S, !second ? &S->ctx_clip : feat ? &S->ctx_clip2f : &S->ctx_clip2);
	ClipState* st = !second ? &S->st_clip : feat ? &S->st_clip2f : &S->st_clip2;
	CHECK(clip_text_encode(st, C, cp, second ? "clip2" : "clip", n_tok, toks, embed, feat, S->clip_skip, !(flags & MLIS_CTEF_NO_NORM)));
	if (weights && embed) {   /* plain per-token scaling of rows 1..n_tok, no renormalisation (mlimgsynth.c:1457-1464) */
		int d = embed->n[0];
		for (int t = 1; t <= n_tok; ++t) for (int i = 0; i < d; ++i) embed->d[(size_t)t * d + i] *= weights[t - 1];
	}
	return 1;
}

int mlis_clip_text_encode(MLIS_Ctx* S, const char* text, MLIS_Tensor* embed, MLIS_Tensor* feat, MLIS_SubModel model, int flags)
{
	int32_t* toks; int n = mlis_text_tokenize(S, text, &toks, model);
	if (n < 0) return n;
	API_TRY(clip_tokens_encode(S, n, toks, NULL, (HTensor*)embed, (HTensor*)feat, model, flags), "mlis_clip_text_encode");
	return 1;
}

/* sinusoidal embedding of scalars (CompVis; cos first) -- mlimgsynth.c:1484-1499 */
static float* sincos_embed(int n, const float* v, int dim, float max_period, float* out)
{
	int half = dim / 2;
	for (int i = 0; i < half; ++i) {
		float freq = exp(-log(max_period) * i / half);
		for (int s = 0; s < n; ++s) { out[s * dim + i] = cos(v[s] * freq); out[s * dim + i + half] = sin(v[s] * freq); }
	}
	return out + (size_t)n * dim;
}

static int text_cond_encode(MLIS_Ctx* S, const PromptText* prompt, HTensor* cond, HTensor* label)
{
	int cte = S->unet_p->clip_norm ? 0 : MLIS_CTEF_NO_NORM;
	int n_tok = prompt_tokenize(S, prompt, S->clip_p);
	if (n_tok < 0) return n_tok;
	CHECK(clip_tokens_encode(S, n_tok, S->tokens, S->tok_w, cond, NULL, MLIS_SUBMODEL_CLIP, cte));
	if (S->unet_p->cond_label) {   /* SDXL: concat both towers, pooled bigG feature + size embeddings (mlimgsynth.c:1520-1558) */
		HTensor e2 = {0}, feat = {0};
		CHECK(clip_tokens_encode(S, n_tok, S->tokens, S->tok_w, &e2, NULL, MLIS_SUBMODEL_CLIP2, cte));
		int n1 = cond->n[0], n2 = e2.n[0], nt = e2.n[1], ne = n1 + n2;
		HTensor cat = {0};
		ht_resize(&cat, ne, nt, 1, 1);
		for (int t = 0; t < nt; ++t) {
			memcpy(cat.d + (size_t)ne * t, cond->d + (size_t)n1 * t, n1 * sizeof(float));
			memcpy(cat.d + (size_t)ne * t + n1, e2.d + (size_t)n2 * t, n2 * sizeof(float));
		}
		ht_free(cond); *cond = cat; ht_free(&e2);
		CHECK(clip_tokens_encode(S, n_tok, S->tokens, NULL, NULL, &feat, MLIS_SUBMODEL_CLIP2, 0));
		ht_resize(label, S->unet_p->ch_adm_in, 1, 1, 1);
		memcpy(label->d, feat.d, n2 * sizeof(float));
		float* ld = label->d + n2;
		float hw[2] = { (float)S->height, (float)S->width }, zz[2] = { 0, 0 };
		ld = sincos_embed(2, hw, 256, 10000, ld);     /* original size */
		ld = sincos_embed(2, zz, 256, 10000, ld);     /* crop top, left */
		ld = sincos_embed(2, hw, 256, 10000, ld);     /* target size */
		ht_free(&feat);
		if (ld != label->d + label->n[0]) FAIL(-1, "label embedding size mismatch");
	}
	return 1;
}

/* ------------------------------------------------------------------ device buffers, image/latent codecs */
static void dev_reserve(float** p, size_t* have, size_t n)
{
	if (n <= *have) return;
	ggml_b200_free(*p); *p = ggml_b200_malloc(n * sizeof(float)); *have = n;
}

static void latent_sync_host(MLIS_Ctx* S)
{
	if (!S->latent_host_stale) return;
	ggml_b200_download(S->latent.d, S->latent_dev, ht_count(&S->latent) * sizeof(float));
	S->latent_host_stale = false;
}
static void image_sync_host(MLIS_Ctx* S)
{
	if (!S->image_host_stale) return;
	ggml_b200_download(S->image.d, S->image_dev, ht_count(&S->image) * sizeof(float));
	S->image_host_stale = false;
}

static int image_finish(MLIS_Ctx* S, int w, int h, int n);

/* latent_dev [lw,lh,4,n] -> image_dev [8lw,8lh,3,n] (+ RGB8 pack), all on the device */
static int decode_dev(MLIS_Ctx* S, int lw, int lh, int n)
{
	int f = S->vae_p->f_down, w = lw * f, h = lh * f;
	size_t per = (size_t)w * h * 3;
	dev_reserve(&S->image_dev, &S->image_dev_n, per * n);
	/* untiled VAE decode: all images of the batch in one graph run, in chunks of <= 2^17 latent pixels (8 images of 512x512) */
	int done_n = 0;
	if (!(S->flags & CF_USE_TAE) && n > 1) {
		int chunk = (1 << 17) / (lw * lh); if (chunk > n) chunk = n;
		while (chunk >= 2 && n - done_n >= chunk) {
			int r = sdvae_decode_batch(&S->st_vdec, graph_ctx_init(S, &S->ctx_vdec), S->vae_p, S->latent_dev + (size_t)done_n * lw * lh * 4, lw, lh, chunk,
				S->image_dev + per * done_n, S->vae_tile);
			if (r < 0) return r;
			if (r == 0) break;
			done_n += chunk;
		}
	}
	for (int i = done_n; i < n; ++i) {
		const float* l = S->latent_dev + (size_t)i * lw * lh * 4;
		float* im = S->image_dev + per * i;
		if (S->flags & CF_USE_TAE) CHECK(sdtae_decode(&S->st_tdec, graph_ctx_init(S, &S->ctx_tdec), S->tae_p, l, lw, lh, im));
		else CHECK(sdvae_decode(&S->st_vdec, graph_ctx_init(S, &S->ctx_vdec), S->vae_p, l, lw, lh, im, S->vae_tile));
	}
	return image_finish(S, w, h, n);
}

/* image_dev [w,h,3,n] in [0,1] -> NaN check, RGB8 pack on the device, ONE synchronising download */
static int image_finish(MLIS_Ctx* S, int w, int h, int n)
{
	size_t per = (size_t)w * h * 3;
	ggml_b200_nonfinite_accumulate(S->image_dev, (int64_t)(per * n));
	/* float -> RGB8: clamp(v*255, 0, 255) truncated (mlimgsynth.c:112-129) */
	if (per * n > S->u8_dev_n) { ggml_b200_free(S->u8_dev); S->u8_dev = ggml_b200_malloc(per * n); S->u8_dev_n = per * n; }
	for (int i = 0; i < n; ++i) ggml_b200_pack_rgb8(S->u8_dev + per * i, S->image_dev + per * i, w, h, 3, 1.0f, 0.0f);
	if (per * n > S->imgex_cap) { ggml_b200_host_free(S->imgex_all); S->imgex_all = ggml_b200_host_malloc(per * n); S->imgex_cap = per * n; }
	ggml_b200_download(S->imgex_all, S->u8_dev, per * n);     /* the one synchronising read of the generation */
	S->img_w = w; S->img_h = h; S->img_n = n;
	ht_resize(&S->image, w, h, 3, n);
	S->image_host_stale = true;
	S->image.flags |= HT_READY;
	if (ggml_b200_nonfinite_check()) FAIL(MLIS_E_NAN, "NaN found in the decoded image");
	return progress(S, MLIS_STAGE_IMAGE_DECODE, 1, 1);
}

/* host image [w,h,3,1] in [0,1] -> latent_dev [w/8,h/8,4] (posterior sample, scaled) */
static int encode_dev(MLIS_Ctx* S, const HTensor* image, float* latent_out_dev)
{
	int w = image->n[0], h = image->n[1], f = S->vae_p->f_down, lw = w / f, lh = h / f;
	size_t n_img = (size_t)w * h * 3, n_lat = (size_t)lw * lh * 4;
	dev_reserve(&S->image_dev, &S->image_dev_n, n_img + n_lat * 3);
	float* img_dev = S->image_dev;
	float* mom_dev = S->image_dev + n_img;          /* [lw,lh,8] */
	float* noise_dev = mom_dev + n_lat * 2;
	ggml_b200_upload(img_dev, image->d, n_img * sizeof(float));
	if (S->flags & CF_USE_TAE) {
		CHECK(sdtae_encode(&S->st_tenc, graph_ctx_init(S, &S->ctx_tenc), S->tae_p, img_dev, w, h, latent_out_dev));
	} else {
		CHECK(sdvae_encode(&S->st_venc, graph_ctx_init(S, &S->ctx_venc), S->vae_p, img_dev, w, h, mom_dev, S->vae_tile));
		/* posterior sample (vae.c:197-220): consumes ONE call of the global noise stream */
		float* noise = xmalloc(n_lat * sizeof(float));
		rng_philox_randn(&g_rng, (unsigned)n_lat, noise);
		ggml_b200_upload(noise_dev, noise, n_lat * sizeof(float));
		free(noise);
		ggml_b200_vae_sample(latent_out_dev, mom_dev, mom_dev + n_lat, noise_dev, S->vae_p->scale_factor, (int64_t)n_lat);
	}
	ggml_b200_nonfinite_accumulate(latent_out_dev, (int64_t)n_lat);
	if (ggml_b200_nonfinite_check()) FAIL(MLIS_E_NAN, "NaN found in encoded latent");
	return progress(S, MLIS_STAGE_IMAGE_ENCODE, 1, 1);
}

int mlis_image_decode(MLIS_Ctx* S, const MLIS_Tensor* latent, MLIS_Tensor* image, int flags)
{
	(void)flags;
	API_TRY(setup(S), "mlis_image_decode");
	if (latent->n[2] != 4) { mlis_err_set("latent wrong shape"); return fail(S, -1, "mlis_image_decode"); }
	int n = latent->n[3], lw = latent->n[0], lh = latent->n[1];
	size_t cnt = (size_t)lw * lh * 4 * n;
	dev_reserve(&S->latent_dev, &S->latent_dev_n, cnt);
	ggml_b200_upload(S->latent_dev, latent->d, cnt * sizeof(float));
	API_TRY(decode_dev(S, lw, lh, n), "mlis_image_decode");
	image_sync_host(S);
	if ((HTensor*)image != &S->image) ht_copy((HTensor*)image, &S->image);
	image->flags |= HT_READY;
	return 1;
}

/* ---- VAE decode tiles spread across GPUs (BASELINE configs[4]); see include/mlimgsynth_b200.h ---- */
int mlis_b200_vae_tile_plan(MLIS_Ctx* S, int lw, int lh, int* n_tiles, int* tile_w_px, int* tile_h_px)
{
	API_TRY(setup(S), "mlis_b200_vae_tile_plan");
	VaeTilePlan T;
	int nt = sdvae_tile_plan(S->vae_p, lw, lh, S->vae_tile, &T);
	if (n_tiles) *n_tiles = nt;
	if (tile_w_px) *tile_w_px = T.n0 * T.f;
	if (tile_h_px) *tile_h_px = T.n1 * T.f;
	return nt;
}

int mlis_b200_vae_tiles_decode(MLIS_Ctx* S, const MLIS_Tensor* latent, int rank, int world, float* tiles_dev)
{
	API_TRY(setup(S), "mlis_b200_vae_tiles_decode");
	if (latent->n[2] != 4 || latent->n[3] != 1) { mlis_err_set("latent wrong shape"); return fail(S, -1, "mlis_b200_vae_tiles_decode"); }
	if (world < 1 || rank < 0 || rank >= world) { mlis_err_set("invalid rank %d of %d", rank, world); return fail(S, -1, "mlis_b200_vae_tiles_decode"); }
	int lw = latent->n[0], lh = latent->n[1];
	size_t cnt = (size_t)lw * lh * 4;
	dev_reserve(&S->latent_dev, &S->latent_dev_n, cnt);
	ggml_b200_upload(S->latent_dev, latent->d, cnt * sizeof(float));
	API_TRY(sdvae_decode_tiles(&S->st_vdec, graph_ctx_init(S, &S->ctx_vdec), S->vae_p, S->latent_dev, lw, lh, S->vae_tile, rank, world, tiles_dev),
		"mlis_b200_vae_tiles_decode");
	ggml_b200_synchronize();       /* the caller hands tiles_dev to its collective on another stream */
	return 1;
}

int mlis_b200_vae_tiles_merge(MLIS_Ctx* S, int lw, int lh, const float* gathered_dev, int world, int slots_per_worker, MLIS_Tensor* image)
{
	API_TRY(setup(S), "mlis_b200_vae_tiles_merge");
	int f = S->vae_p->f_down, w = lw * f, h = lh * f;
	dev_reserve(&S->image_dev, &S->image_dev_n, (size_t)w * h * 3);
	API_TRY(sdvae_merge_tiles(S->vae_p, lw, lh, S->vae_tile, gathered_dev, world, slots_per_worker, S->image_dev), "mlis_b200_vae_tiles_merge");
	API_TRY(image_finish(S, w, h, 1), "mlis_b200_vae_tiles_merge");
	if (image) { image_sync_host(S); if ((HTensor*)image != &S->image) ht_copy((HTensor*)image, &S->image); image->flags |= HT_READY; }
	return 1;
}

int mlis_image_encode(MLIS_Ctx* S, const MLIS_Tensor* image, MLIS_Tensor* latent, int flags)
{
	(void)flags;
	API_TRY(setup(S), "mlis_image_encode");
	int f = S->vae_p->f_down, lw = image->n[0] / f, lh = image->n[1] / f;
	if (image->n[0] % f || image->n[1] % f || image->n[2] != 3 || image->n[3] != 1) { mlis_err_set("invalid input image shape"); return fail(S, -1, "mlis_image_encode"); }
	dev_reserve(&S->latent_dev, &S->latent_dev_n, (size_t)lw * lh * 4);
	API_TRY(encode_dev(S, (const HTensor*)image, S->latent_dev), "mlis_image_encode");
	ht_resize((HTensor*)latent, lw, lh, 4, 1);
	ggml_b200_download(latent->d, S->latent_dev, (size_t)lw * lh * 4 * sizeof(float));
	return 1;
}

int mlis_mask_encode(MLIS_Ctx* S, const MLIS_Tensor* mask, MLIS_Tensor* lmask, int flags)
{
	(void)flags;
	int f = S->vae_p ? S->vae_p->f_down : 8;   /* box average f x f (localtensor.c:161-194) */
	int w = mask->n[0], h = mask->n[1], ow = w / f, oh = h / f;
	HTensor out = {0};
	ht_resize(&out, ow, oh, 1, 1);
	for (int y = 0; y < oh; ++y) for (int x = 0; x < ow; ++x) {
		float s = 0;
		for (int dy = 0; dy < f; ++dy) for (int dx = 0; dx < f; ++dx) s += mask->d[(size_t)(y * f + dy) * w + x * f + dx];
		out.d[(size_t)y * ow + x] = s / (f * f);
	}
	ht_free((HTensor*)lmask); *(HTensor*)lmask = out;
	return 1;
}

/* ------------------------------------------------------------------ denoising */
struct dxdt_args { MLIS_Ctx* S; };

/* dx/dt for the solvers: ONE batched UNet evaluation (cond | uncond), then a single fused kernel for
 * the v-prediction mix (unet.c:490-494) and the CFG combine (mlimgsynth.c:1580-1583):
 *   dx = f (c_out o_c + c_skip x) + (1-f) (c_out o_u + c_skip x) */
static int denoise_dxdt(Solver* sol, float t, const float* x, float* dx)
{
	if (!(t >= 0)) return 0;
	MLIS_Ctx* S = ((struct dxdt_args*)sol->user)->S;
	const float* out;
	CHECK(unet_denoise_run(&S->unet, x, t, &out));
	int64_t n1 = sol->n;
	ggml_b200_nonfinite_accumulate(out, n1 * S->unet.n_rep);
	float c_out = 1, c_skip = 0, f = S->cfg_scale;
	if (S->unet_p->vparam) { c_skip = t / (t * t + 1); c_out = 1 / sqrt(t * t + 1); }   /* reference's mixed float/double forms */
	float* outs[1] = { dx };
	if (S->cfg_half >= 0 && S->cfg_scale > 1) {
		/* cross-GPU CFG split: this rank evaluated ONE half (its conditioning rows); the peer's half arrives through the
		   caller's exchange (a 2-rank all-gather of n1 floats per evaluation), then the same combine as below */
		dev_reserve(&S->cfg_other_dev, &S->cfg_other_n, (size_t)n1);
		ggml_b200_synchronize();      /* the exchange runs on the caller's stream */
		int r = S->cfg_exchange(S->cfg_exchange_user, out, S->cfg_other_dev, (size_t)n1);
		if (r < 0) FAIL(r, "CFG exchange callback failed");
		const float* oc = S->cfg_half == 0 ? out : S->cfg_other_dev;
		const float* ou = S->cfg_half == 0 ? S->cfg_other_dev : out;
		const float* ins[3] = { oc, ou, x };
		float c[3] = { c_out * f, c_out * (1 - f), c_skip };
		ggml_b200_lincomb(1, outs, c_skip != 0 ? 3 : 2, ins, c, n1);
	} else if (S->unet.n_rep == 2) {
		const float* ins[3] = { out, out + n1, x };
		float c[3] = { c_out * f, c_out * (1 - f), c_skip };
		ggml_b200_lincomb(1, outs, c_skip != 0 ? 3 : 2, ins, c, n1);
	} else {
		const float* ins[2] = { out, x };
		float c[2] = { c_out, c_skip };
		ggml_b200_lincomb(1, outs, c_skip != 0 ? 2 : 1, ins, c, n1);
	}
	return 1;
}

static void infotext_update(MLIS_Ctx* S, int w, int h)
{
	char b[4096]; int n = 0;
#define AP(...) do { if (n < (int)sizeof(b)) n += snprintf(b + n, sizeof(b) - n, __VA_ARGS__); } while (0)
	AP("%s\n", S->prompt_raw ? S->prompt_raw : "");
	if (S->nprompt_raw && *S->nprompt_raw) AP("Negative prompt: %s\n", S->nprompt_raw);
	AP("Seed: %llu, Sampler: %s", (unsigned long long)g_rng.seed, mlis_method_str(S->sampler.c.method));
	if (S->sampler.c.s_ancestral == 1) AP(" ancestral");
	AP(", Schedule type: %s", mlis_sched_str(S->sampler.c.sched));
	if (S->sampler.c.s_ancestral > 0) AP(", Ancestral: %g", S->sampler.c.s_ancestral);
	if (S->sampler.c.s_noise > 0) AP(", SNoise: %g", S->sampler.c.s_noise);
	if (S->cfg_scale > 1) AP(", CFG scale: %g", S->cfg_scale);
	if (S->sampler.c.f_t_ini < 1) AP(", Mode: %s, f_t_ini: %g", S->sampler.c.lmask_dev ? "inpaint" : "img2img", S->sampler.c.f_t_ini);
	AP(", Steps: %d, NFE: %d, Size: %dx%d, Clip skip: %d", S->sampler.n_step, S->prg.nfe, w, h, S->clip_skip);
	const char* base = S->path_model ? strrchr(S->path_model, '/') : NULL;
	base = base ? base + 1 : S->path_model ? S->path_model : "";
	const char* ext = strrchr(base, '.');
	AP(", Model: %.*s", (int)(ext ? ext - base : (long)strlen(base)), base);
	if (S->flags & CF_USE_TAE) AP(", VAE: tae");
	if (S->n_batch > 1) AP(", Batch size: %d", S->n_batch);
	AP(", Version: MLImgSynth-B200 v%s", MLIS_VERSION_STR);
#undef AP
	set_str(&S->infotext, b);
}

static void prompt_clear(MLIS_Ctx* S)   /* options cleared after each generation (mlimgsynth.c:696-709) */
{
	prompt_text_clear(&S->prompt); prompt_text_clear(&S->nprompt);
	set_str(&S->prompt_raw, NULL); set_str(&S->nprompt_raw, NULL);
	S->sampler.c.f_t_ini = 1; S->sampler.c.f_t_end = 0;
	S->tuflags = 0;
	loras_free(S, true);
}

static int generate(MLIS_Ctx* S)
{
	CHECK(setup(S));
	const UnetParams* P = S->unet_p;
	int nb = S->n_batch > 0 ? S->n_batch : 1;
	S->t_last = time_now(); memset(&S->prg, 0, sizeof(S->prg));
	double t_start = S->t_last;
	int f = S->vae_p->f_down, w = S->width / f, h = S->height / f;

	/* initial latent: encoded input image (img2img), caller-provided latent, or zeros */
	bool have_latent = false;
	if (S->tuflags & MLIS_TUF_IMAGE) {
		if (nb > 1) FAIL(-1, "img2img supports batch size 1");
		w = S->image.n[0] / f; h = S->image.n[1] / f;
		dev_reserve(&S->latent_dev, &S->latent_dev_n, (size_t)w * h * 4);
		HTensor img = S->image; img.n[3] = 1;
		CHECK(encode_dev(S, &img, S->latent_dev));
		have_latent = true;
	} else if (S->tuflags & MLIS_TUF_LATENT) {
		latent_sync_host(S);
		w = S->latent.n[0]; h = S->latent.n[1];
		if (S->latent.n[2] != P->n_ch_in || S->latent.n[3] != nb) FAIL(-1, "input latent has the wrong shape");
		dev_reserve(&S->latent_dev, &S->latent_dev_n, ht_count(&S->latent));
		ggml_b200_upload(S->latent_dev, S->latent.d, ht_count(&S->latent) * sizeof(float));
		have_latent = true;
	}
	int64_t n_per = (int64_t)w * h * P->n_ch_in, n_all = n_per * nb;
	dev_reserve(&S->latent_dev, &S->latent_dev_n, (size_t)n_all);
	if (!have_latent) ggml_b200_memset(S->latent_dev, 0, (size_t)n_all * sizeof(float));
	ht_resize(&S->latent, w, h, P->n_ch_in, nb);
	int w_img = w * f, h_img = h * f;
	log_info("Output size: %dx%d x%d", w_img, h_img, nb);

	/* inpainting mask -> latent mask (box average), kept on the device */
	if (S->tuflags & MLIS_TUF_MASK) { mlis_mask_encode(S, (MLIS_Tensor*)&S->mask, (MLIS_Tensor*)&S->lmask, 0); S->tuflags |= MLIS_TUF_LMASK; }
	const float* lmask_dev = NULL;
	if ((S->tuflags & MLIS_TUF_LMASK) && S->lmask.d) {
		if (S->lmask.n[0] != w || S->lmask.n[1] != h) FAIL(-1, "latent mask has the wrong shape");
		dev_reserve(&S->lmask_dev, &S->lmask_dev_n, (size_t)w * h);
		ggml_b200_upload(S->lmask_dev, S->lmask.d, (size_t)w * h * sizeof(float));
		lmask_dev = S->lmask_dev;
		log_info("In-painting with mask");
	}

	/* conditioning */
	bool cfg = S->cfg_scale > 1;
	if (!(S->tuflags & MLIS_TUF_CONDITIONING)) {
		CHECK(text_cond_encode(S, &S->prompt, &S->cond, &S->label));
		if (cfg) {
			CHECK(text_cond_encode(S, &S->nprompt, &S->ncond, &S->nlabel));
			if (P->uncond_empty_zero && !(S->nprompt_raw && *S->nprompt_raw)) memset(S->ncond.d, 0, ht_count(&S->ncond) * sizeof(float));
		}
		CHECK(progress(S, MLIS_STAGE_COND_ENCODE, 1, 1));
	}
	S->image.flags &= ~HT_READY;

	/* sampler + batched UNet (images x CFG halves) */
	const bool split = cfg && S->cfg_half >= 0;      /* the other CFG half runs on a peer GPU */
	int n_rep = (cfg && !split) ? 2 : 1;
	S->sampler.unet_p = P;
	S->sampler.nfe_per_dxdt = cfg ? 2 : 1;
	S->sampler.c.lmask_dev = lmask_dev; S->sampler.c.mask_pix = (int64_t)w * h;
	for (int i = 0; i < nb; ++i) { S->rngs[i].seed = g_rng.seed + i; S->rngs[i].offset = g_rng.offset; }
	S->sampler.rng = S->rngs; S->sampler.n_rng = nb; S->sampler.n_per_image = n_per;
	struct dxdt_args A = { S };
	S->sampler.solver.dxdt = denoise_dxdt; S->sampler.solver.user = &A;
	CHECK(dnsamp_init(&S->sampler));
	CHECK(unet_denoise_init(&S->unet, graph_ctx_init(S, &S->ctx_unet), P, w, h, nb, n_rep));
	{   /* conditioning rows for the graph batch: [cond x nb | uncond x nb] */
		size_t nc = ht_count(&S->cond), nl = P->ch_adm_in;
		float* cb = xmalloc(nc * nb * n_rep * sizeof(float));
		float* lb = nl ? xmalloc(nl * nb * n_rep * sizeof(float)) : NULL;
		for (int r = 0; r < n_rep; ++r) for (int i = 0; i < nb; ++i) {
			const bool neg = split ? S->cfg_half == 1 : r == 1;
			const HTensor *c = neg ? &S->ncond : &S->cond, *l = neg ? &S->nlabel : &S->label;
			if (ht_count(c) != nc) { free(cb); free(lb); FAIL(-1, "conditioning tensors have different shapes"); }
			memcpy(cb + nc * (r * nb + i), c->d, nc * sizeof(float));
			if (lb) memcpy(lb + nl * (r * nb + i), l->d, nl * sizeof(float));
		}
		int rr = unet_cond_set(&S->unet, cb, lb);
		ggml_b200_synchronize();
		free(cb); free(lb);
		CHECK(rr);
	}
	log_info("Generating (solver: %s, sched: %s, ancestral: %g, snoise: %g, cfg-s: %g, steps: %d, nfe/s: %d, batch: %d)",
		mlis_method_str(S->sampler.c.method), mlis_sched_str(S->sampler.c.sched), S->sampler.c.s_ancestral, S->sampler.c.s_noise,
		S->cfg_scale, S->sampler.n_step, S->sampler.nfe_per_step, nb);

	int r;
	while ((r = dnsamp_step(&S->sampler, S->latent_dev)) > 0) {
		S->prg.nfe = S->unet.nfe;
		S->latent_host_stale = true;
		CHECK(progress(S, MLIS_STAGE_DENOISE, S->sampler.i_step, S->sampler.n_step));
	}
	CHECK(r);
	g_rng.offset = S->rngs[0].offset;     /* batch 1 leaves the global stream exactly where the reference would */
	S->latent_host_stale = true;

	if (!(S->flags & CF_NO_DECODE)) CHECK(decode_dev(S, w, h, nb));
	else if (ggml_b200_nonfinite_check()) FAIL(MLIS_E_NAN, "NaN found in UNet output");
	infotext_update(S, w_img, h_img);
	prompt_clear(S);
	log_info("Generation done {%.3fs}", time_now() - t_start);
	return 1;
}

int mlis_generate(MLIS_Ctx* S)
{
	API_TRY(generate(S), "mlis_generate");
	return 1;
}

MLIS_Image* mlis_image_get(MLIS_Ctx* S, int idx)
{
	if (!(S->image.flags & HT_READY) || !S->imgex_all) { snprintf(S->errstr, sizeof(S->errstr), "image not ready"); return NULL; }
	if (idx < 0 || idx >= S->img_n) { snprintf(S->errstr, sizeof(S->errstr), "image index %d out of range (batch %d)", idx, S->img_n); return NULL; }
	size_t per = (size_t)S->img_w * S->img_h * 3;
	S->imgex.d = S->imgex_all + per * idx; S->imgex.sz = per; S->imgex.w = S->img_w; S->imgex.h = S->img_h; S->imgex.c = 3; S->imgex.flags = 0;
	return &S->imgex;
}

/* Cross-GPU CFG split (opt-in): see include/mlimgsynth_b200.h */
int mlis_b200_cfg_split_set(MLIS_Ctx* S, int half, MLIS_B200_CfgExchange fn, void* user)
{
	if (half >= 0 && (half > 1 || !fn)) { snprintf(S->errstr, sizeof(S->errstr), "cfg split: half must be 0 or 1 and needs an exchange callback"); return MLIS_E_OPT_VALUE; }
	S->cfg_half = half < 0 ? -1 : half; S->cfg_exchange = fn; S->cfg_exchange_user = user;
	return 1;
}

/* The RGB8 images of the last generation / decode as the pack kernel left them in HBM ([n][h][w][3] bytes): multi-GPU callers
 * gather them device-to-device instead of re-uploading the host copies. Valid until the next generate / decode call. */
int mlis_b200_images_device(MLIS_Ctx* S, const uint8_t** dev, int* w, int* h, int* n)
{
	if (!(S->image.flags & HT_READY) || !S->u8_dev) { snprintf(S->errstr, sizeof(S->errstr), "image not ready"); return -1; }
	if (dev) *dev = S->u8_dev;
	if (w) *w = S->img_w;
	if (h) *h = S->img_h;
	if (n) *n = S->img_n;
	return 1;
}

const char* mlis_infotext_get(MLIS_Ctx* S, int idx) { (void)idx; return S->infotext; }

MLIS_Tensor* mlis_tensor_get(MLIS_Ctx* S, MLIS_TensorId id)
{
	switch (id) {
	case MLIS_TENSOR_IMAGE: if (S->image.d) image_sync_host(S); return (MLIS_Tensor*)&S->image;
	case MLIS_TENSOR_MASK: return (MLIS_Tensor*)&S->mask;
	case MLIS_TENSOR_LATENT: if (S->latent.d) latent_sync_host(S); return (MLIS_Tensor*)&S->latent;
	case MLIS_TENSOR_LMASK: return (MLIS_Tensor*)&S->lmask;
	case MLIS_TENSOR_COND: return (MLIS_Tensor*)&S->cond;
	case MLIS_TENSOR_LABEL: return (MLIS_Tensor*)&S->label;
	case MLIS_TENSOR_NCOND: return (MLIS_Tensor*)&S->ncond;
	case MLIS_TENSOR_NLABEL: return (MLIS_Tensor*)&S->nlabel;
	default:
		if ((unsigned)id >= MLIS_TENSOR_TMP && (unsigned)id < MLIS_TENSOR_TMP + 8) return (MLIS_Tensor*)&S->tmp[id - MLIS_TENSOR_TMP];
		return NULL;
	}
}

int mlis_unet_eval(MLIS_Ctx* S, const MLIS_Tensor* x, const MLIS_Tensor* cond, const MLIS_Tensor* label, float sigma, MLIS_Tensor* dx)
{
	API_TRY(setup(S), "mlis_unet_eval");
	const UnetParams* P = S->unet_p;
	int w = x->n[0], h = x->n[1], nb = x->n[3];
	size_t n = (size_t)w * h * P->n_ch_in * nb;
	API_TRY(unet_denoise_init(&S->unet, graph_ctx_init(S, &S->ctx_unet), P, w, h, nb, 1), "mlis_unet_eval");
	API_TRY(unet_cond_set(&S->unet, cond->d, label ? label->d : NULL), "mlis_unet_eval");
	dev_reserve(&S->latent_dev, &S->latent_dev_n, n * 2);
	ggml_b200_upload(S->latent_dev, x->d, n * sizeof(float));
	const float* out;
	API_TRY(unet_denoise_run(&S->unet, S->latent_dev, sigma, &out), "mlis_unet_eval");
	float* dxd = S->latent_dev + n;
	float c_out = 1, c_skip = 0;
	if (P->vparam) { c_skip = sigma / (sigma * sigma + 1); c_out = 1 / sqrt(sigma * sigma + 1); }
	float* outs[1] = { dxd }; const float* ins[2] = { out, S->latent_dev }; float c[2] = { c_out, c_skip };
	ggml_b200_lincomb(1, outs, c_skip != 0 ? 2 : 1, ins, c, (int64_t)n);
	ht_resize((HTensor*)dx, w, h, P->n_ch_in, nb);
	ggml_b200_download(dx->d, dxd, n * sizeof(float));
	return 1;
}

/* ------------------------------------------------------------------ tensor helpers */
void   mlis_tensor_free(MLIS_Tensor* T) { ht_free((HTensor*)T); }
size_t mlis_tensor_count(const MLIS_Tensor* T) { return ht_count((const HTensor*)T); }
void   mlis_tensor_resize(MLIS_Tensor* T, int n0, int n1, int n2, int n3) { ht_resize((HTensor*)T, n0, n1, n2, n3); }
void   mlis_tensor_resize_like(MLIS_Tensor* T, const MLIS_Tensor* O) { ht_resize((HTensor*)T, O->n[0], O->n[1], O->n[2], O->n[3]); }
void   mlis_tensor_copy(MLIS_Tensor* T, const MLIS_Tensor* O) { ht_copy((HTensor*)T, (const HTensor*)O); }
float  mlis_tensor_similarity(const MLIS_Tensor* A, const MLIS_Tensor* B)
{
	size_t n = mlis_tensor_count(A);
	if (n != mlis_tensor_count(B)) return 0;
	double ab = 0, aa = 0, bb = 0;
	for (size_t i = 0; i < n; ++i) { ab += (double)A->d[i] * B->d[i]; aa += (double)A->d[i] * A->d[i]; bb += (double)B->d[i] * B->d[i]; }
	return (float)(ab / sqrt(aa * bb));
}
