#include "tstore.h"
#include <ctype.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

static uint32_t str_hash(const char* s) { uint32_t h = 2166136261u; for (; *s; ++s) h = (h ^ (uint8_t)*s) * 16777619u; return h; }

static void hash_rebuild(TStore* S)
{
	S->hash_cap = 64; while (S->hash_cap < S->n * 2 + 16) S->hash_cap *= 2;
	free(S->hash); S->hash = xmalloc(S->hash_cap * sizeof(int));
	for (int i = 0; i < S->hash_cap; ++i) S->hash[i] = -1;
	for (int i = 0; i < S->n; ++i) {
		uint32_t h = str_hash(S->e[i].key) & (S->hash_cap - 1);
		while (S->hash[h] >= 0) h = (h + 1) & (S->hash_cap - 1);
		S->hash[h] = i;
	}
}

TSEntry* tstore_find(const TStore* S, const char* key)
{
	if (!S->hash) return NULL;
	uint32_t h = str_hash(key) & (S->hash_cap - 1);
	while (S->hash[h] >= 0) {
		if (!strcmp(S->e[S->hash[h]].key, key)) return &S->e[S->hash[h]];
		h = (h + 1) & (S->hash_cap - 1);
	}
	return NULL;
}

int64_t tsentry_count(const TSEntry* e) { int64_t n = 1; for (int i = 0; i < e->ndim; ++i) n *= e->shape[i]; return n; }

TSEntry* tstore_add(TStore* S, const char* key, int dtype, int ndim, const int64_t* shape, const uint8_t* data, size_t nbytes, bool owned)
{
	TSEntry* old = tstore_find(S, key);
	TSEntry ne = { xstrdup(key), dtype, ndim, {1, 1, 1, 1}, data, nbytes, owned };
	for (int i = 0; i < ndim && i < 4; ++i) ne.shape[i] = shape[i];
	if (old) {   /* replace in place (LoRA merge result) */
		free(old->key); if (old->owned) free((void*)old->data);
		*old = ne;
		return old;
	}
	ARR_PUSH(S->e, S->n, S->cap, ne);
	if (S->n * 2 + 16 > S->hash_cap) hash_rebuild(S);
	else {
		uint32_t h = str_hash(key) & (S->hash_cap - 1);
		while (S->hash[h] >= 0) h = (h + 1) & (S->hash_cap - 1);
		S->hash[h] = S->n - 1;
	}
	return &S->e[S->n - 1];
}

void tstore_free(TStore* S)
{
	for (int i = 0; i < S->n; ++i) { free(S->e[i].key); if (S->e[i].owned) free((void*)S->e[i].data); }
	free(S->e); free(S->hash);
	for (int i = 0; i < S->n_map; ++i) munmap(S->maps[i].p, S->maps[i].size);
	free(S->maps);
	memset(S, 0, sizeof(*S));
}

/* ---- minimal JSON scanner for the safetensors header ---- */
typedef struct { const char* p; const char* end; } JS;
static void js_ws(JS* j) { while (j->p < j->end && isspace((unsigned char)*j->p)) j->p++; }
static bool js_ch(JS* j, char c) { js_ws(j); if (j->p < j->end && *j->p == c) { j->p++; return true; } return false; }
static bool js_str(JS* j, char* out, size_t n)
{
	js_ws(j);
	if (j->p >= j->end || *j->p != '"') return false;
	j->p++;
	size_t k = 0;
	while (j->p < j->end && *j->p != '"') {
		char c = *j->p++;
		if (c == '\\' && j->p < j->end) c = *j->p++;
		if (k + 1 < n) out[k++] = c;
	}
	if (j->p >= j->end) return false;
	j->p++; out[k] = 0;
	return true;
}
static bool js_int(JS* j, int64_t* v)
{
	js_ws(j);
	char* e; long long x = strtoll(j->p, &e, 10);
	if (e == j->p) return false;
	j->p = e; *v = x; return true;
}
static bool js_skip(JS* j)   /* skip any value */
{
	js_ws(j);
	if (j->p >= j->end) return false;
	if (*j->p == '"') { char tmp[8]; return js_str(j, tmp, sizeof(tmp)); }
	if (*j->p == '{' || *j->p == '[') {
		char open = *j->p, close = open == '{' ? '}' : ']';
		int depth = 0; bool in_str = false;
		for (; j->p < j->end; j->p++) {
			char c = *j->p;
			if (in_str) { if (c == '\\') j->p++; else if (c == '"') in_str = false; continue; }
			if (c == '"') in_str = true;
			else if (c == open) depth++;
			else if (c == close && --depth == 0) { j->p++; return true; }
		}
		return false;
	}
	while (j->p < j->end && *j->p != ',' && *j->p != '}' && *j->p != ']') j->p++;
	return true;
}

static size_t dtype_size(int dtype) { return dtype == TS_F32 ? 4 : (dtype == TS_F16 || dtype == TS_BF16) ? 2 : 0; }

/* Parses the header of a mapped file into `S`. Nothing from the file is trusted: every entry must lie inside the data
 * section and its byte range must be exactly shape x element size, so consumers may read tsentry_count() elements. */
static int safetensors_index(TStore* S, const uint8_t* base, uint64_t file_size, uint64_t hlen, ts_name_conv conv, const char* prefix)
{
	const uint8_t* data0 = base + 8 + hlen;
	const uint64_t data_size = file_size - 8 - hlen;
	JS j = { (const char*)base + 8, (const char*)base + 8 + hlen };
	if (!js_ch(&j, '{')) FAIL(-1, "safetensors header: expected object");
	char key[256], conv_key[256], full[320], field[32], dt[16];
	while (!js_ch(&j, '}')) {
		js_ch(&j, ',');
		if (!js_str(&j, key, sizeof(key)) || !js_ch(&j, ':')) FAIL(-1, "safetensors header: bad key");
		if (!strcmp(key, "__metadata__")) { if (!js_skip(&j)) FAIL(-1, "safetensors header: bad metadata"); continue; }
		if (!js_ch(&j, '{')) FAIL(-1, "safetensors header: bad entry '%s'", key);
		int64_t shape[8], off0 = -1, off1 = -1; int nd = 0; dt[0] = 0;
		while (!js_ch(&j, '}')) {
			js_ch(&j, ',');
			if (!js_str(&j, field, sizeof(field)) || !js_ch(&j, ':')) FAIL(-1, "safetensors header: bad field in '%s'", key);
			if (!strcmp(field, "dtype")) { if (!js_str(&j, dt, sizeof(dt))) FAIL(-1, "bad dtype"); }
			else if (!strcmp(field, "shape")) {
				if (!js_ch(&j, '[')) FAIL(-1, "bad shape");
				while (!js_ch(&j, ']')) { js_ch(&j, ','); if (nd >= 8 || !js_int(&j, &shape[nd++])) FAIL(-1, "bad shape in '%s'", key); }
			}
			else if (!strcmp(field, "data_offsets")) {
				if (!js_ch(&j, '[') || !js_int(&j, &off0) || !js_ch(&j, ',') || !js_int(&j, &off1) || !js_ch(&j, ']')) FAIL(-1, "bad offsets");
			}
			else if (!js_skip(&j)) FAIL(-1, "bad field value");
		}
		int r = 1;
		if (conv) { r = conv(key, conv_key, sizeof(conv_key)); if (r <= 0) continue; }
		else snprintf(conv_key, sizeof(conv_key), "%s", key);
		snprintf(full, sizeof(full), "%s%s", prefix ? prefix : "", conv_key);
		int dtype = !strcmp(dt, "F32") ? TS_F32 : !strcmp(dt, "F16") ? TS_F16 : !strcmp(dt, "BF16") ? TS_BF16 : TS_OTHER;
		if (nd > 4 || dtype == TS_OTHER) continue;      /* not consumable by the engine (no quantised / integer weights on this path) */
		int64_t rs[4] = {1, 1, 1, 1}, count = 1;
		for (int i = 0; i < nd; ++i) {
			rs[i] = shape[nd - 1 - i];
			if (rs[i] < 0 || (rs[i] && count > INT64_MAX / 8 / rs[i])) FAIL(-1, "tensor '%s': invalid shape", key);
			count *= rs[i];
		}
		if (off0 < 0 || off1 < off0 || (uint64_t)off1 > data_size) FAIL(-1, "tensor '%s' data out of file bounds", key);
		if ((uint64_t)(off1 - off0) != (uint64_t)count * dtype_size(dtype))
			FAIL(-1, "tensor '%s': %lld bytes in the file, shape x dtype needs %lld", key, (long long)(off1 - off0), (long long)(count * (int64_t)dtype_size(dtype)));
		if (r == 2) {
			/* fused OpenCLIP attn.in_proj_{weight,bias}: split the outer dim in q,k,v thirds (mlimgsynth.c:990-1030) */
			const char* tail = strstr(full, "in_proj_");
			if (!tail || nd < 1 || rs[nd - 1] % 3) FAIL(-1, "cannot split fused projection '%s'", key);
			bool is_w = !strcmp(tail, "in_proj_weight");
			int64_t part[4] = { rs[0], rs[1], rs[2], rs[3] };
			part[nd - 1] /= 3;
			size_t pb = (size_t)(off1 - off0) / 3;
			static const char* names[3] = { "q_proj", "k_proj", "v_proj" };
			for (int k = 0; k < 3; ++k) {
				char nm[320];
				snprintf(nm, sizeof(nm), "%.*s%s.%s", (int)(tail - full), full, names[k], is_w ? "weight" : "bias");
				tstore_add(S, nm, dtype, nd, part, data0 + off0 + pb * k, pb, false);
			}
			continue;
		}
		tstore_add(S, full, dtype, nd, rs, data0 + off0, (size_t)(off1 - off0), false);
	}
	return S->n;
}

int tstore_read_safetensors(TStore* S, const char* path, ts_name_conv conv, const char* prefix)
{
	int fd = open(path, O_RDONLY);
	if (fd < 0) FAIL(-6, "could not open '%s'", path);
	struct stat st;
	if (fstat(fd, &st) < 0 || st.st_size < 8) { close(fd); FAIL(-1, "'%s' is not a safetensors file", path); }
	void* map = mmap(NULL, st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
	close(fd);
	if (map == MAP_FAILED) FAIL(-1, "could not map '%s'", path);
	uint64_t hlen; memcpy(&hlen, map, 8);
	if (hlen > (uint64_t)st.st_size - 8) { munmap(map, st.st_size); FAIL(-1, "'%s' is not a safetensors file", path); }
	int n_before = S->n;
	int r = safetensors_index(S, map, (uint64_t)st.st_size, hlen, conv, prefix);
	if (r < 0 && S->n == n_before) { munmap(map, st.st_size); return r; }    /* nothing references the mapping */
	/* a store may hold several files (model + TAE): every mapping lives until tstore_free */
	struct TSMap m = { map, (size_t)st.st_size };
	ARR_PUSH(S->maps, S->n_map, S->cap_map, m);
	return r;
}

/* ---- dtype conversion on the host (tensorstore.c:185-230 role) ---- */
static inline float f16_to_f32(uint16_t h) { _Float16 x; memcpy(&x, &h, 2); return (float)x; }
static inline uint16_t f32_to_f16(float f) { _Float16 x = (_Float16)f; uint16_t h; memcpy(&h, &x, 2); return h; }
static inline float bf16_to_f32(uint16_t h) { uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; }

const void* tsentry_as(const TSEntry* e, int want, void** to_free)
{
	*to_free = NULL;
	if (e->dtype == want) return e->data;
	int64_t n = tsentry_count(e);
	if (want == TS_F32) {
		float* o = xmalloc(n * 4);
		const uint16_t* s = (const uint16_t*)e->data;
		if (e->dtype == TS_F16) for (int64_t i = 0; i < n; ++i) o[i] = f16_to_f32(s[i]);
		else if (e->dtype == TS_BF16) for (int64_t i = 0; i < n; ++i) o[i] = bf16_to_f32(s[i]);
		else { free(o); return NULL; }
		*to_free = o; return o;
	}
	if (want == TS_F16) {
		uint16_t* o = xmalloc(n * 2);
		if (e->dtype == TS_F32) { const float* s = (const float*)e->data; for (int64_t i = 0; i < n; ++i) o[i] = f32_to_f16(s[i]); }
		else if (e->dtype == TS_BF16) { const uint16_t* s = (const uint16_t*)e->data; for (int64_t i = 0; i < n; ++i) o[i] = f32_to_f16(bf16_to_f32(s[i])); }
		else { free(o); return NULL; }
		*to_free = o; return o;
	}
	return NULL;
}
