/* tstore.h -- weight files: mmap'd safetensors index + name mapping to the module-graph paths.
 * Role of the reference's ccompute/tensorstore*.c + tensor_name_conv.c on this path:
 * only what the B200 engine consumes (F32/F16/BF16 safetensors), no GGUF, no quantised types. */
#pragma once
#include "base.h"

enum { TS_F32 = 0, TS_F16 = 1, TS_BF16 = 2, TS_OTHER = 9 };

typedef struct TSEntry {
	char*    key;          /* internal (module-graph) name, e.g. unet.in.1.0.conv1.weight */
	int      dtype;        /* TS_* */
	int      ndim;
	int64_t  shape[4];     /* ggml order (innermost first): reversed file shape (tensorstore_safet.c:138-142) */
	const uint8_t* data;   /* inside the mapping (or owned when `owned`) */
	size_t   nbytes;
	bool     owned;
} TSEntry;

typedef struct TStore {
	TSEntry* e; int n, cap;
	int*     hash; int hash_cap;       /* open addressing over key strings */
	struct TSMap { void* p; size_t size; } *maps; int n_map, cap_map;   /* one mapping per file read into the store */
} TStore;

/* Reads the index of a safetensors file. Every key is passed through `conv` (NULL: identity);
 * keys it rejects are dropped (as the reference does, mlimgsynth.c:1040-1044). `prefix` is
 * prepended to converted names (TAE files: "tae."). */
typedef int (*ts_name_conv)(const char* file_key, char* out, size_t out_sz);
int  tstore_read_safetensors(TStore* S, const char* path, ts_name_conv conv, const char* prefix);
void tstore_free(TStore* S);
TSEntry* tstore_find(const TStore* S, const char* key);
int64_t  tsentry_count(const TSEntry* e);
/* Adds an entry that aliases/owns memory (fused-QKV split, LoRA-merged weights). */
TSEntry* tstore_add(TStore* S, const char* key, int dtype, int ndim, const int64_t* shape, const uint8_t* data, size_t nbytes, bool owned);
/* Returns the tensor converted to f16 or f32 in a malloc'd buffer (or NULL + *direct when the file
 * bytes can be used as they are). */
const void* tsentry_as(const TSEntry* e, int want_dtype, void** to_free);

/* CompVis/LDM, OpenCLIP and HF-CLIP checkpoint names -> module-graph names
 * (behaviour of tensor_name_conv.c:274 tnconv_sd for the non-diffusers layouts).
 * Returns 1 mapped, 2 mapped fused in_proj (to be split in thirds), 0 unknown. */
int tnconv_sd(const char* file_key, char* out, size_t out_sz);
