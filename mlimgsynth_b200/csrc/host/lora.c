#include "lora.h"
#include "ggml-b200.h"

int lora_name_conv(const char* key, char* out, size_t out_sz)
{
	if (strncmp(key, "lora_", 5)) return 0;
	key += 5;
	const char* dot = strchr(key, '.');        /* module path ends at the first '.', the rest is the LoRA field */
	if (!dot) return 0;
	char mod[256], conv[256];
	size_t n = (size_t)(dot - key);
	if (n + 2 > sizeof(mod)) return 0;
	memcpy(mod, key, n); mod[n] = '.'; mod[n + 1] = 0;   /* trailing separator so "..._to_out_0" matches "to_out.0." */
	if (tnconv_sd(mod, conv, sizeof(conv)) <= 0) return 0;
	size_t cl = strlen(conv);
	if (cl && conv[cl - 1] == '.') conv[cl - 1] = 0;
	snprintf(out, out_sz, "%s%s", conv, dot);
	return 1;
}

static float scalar_of(const TSEntry* e)
{
	void* tmp; const float* p = tsentry_as(e, TS_F32, &tmp);
	float v = p ? p[0] : 0;
	free(tmp);
	return v;
}

/* One tensor: W <- W + scale * up . down in the weight type `ts_dt` (TS_F16 / TS_F32), on the device. */
static int lora_apply_one(TStore* dst, TSEntry* w, const TSEntry* ld, const TSEntry* lu, float scale, int ts_dt)
{
	int64_t r = ld->shape[ld->ndim - 1], n0 = tsentry_count(ld) / r, n1 = tsentry_count(lu) / r;
	const size_t es = ts_dt == TS_F32 ? 4 : 2;
	void *t0 = NULL, *t1 = NULL, *t2 = NULL; char* dev = NULL; uint8_t* merged = NULL;
	int R = 1;
	const void* hw = tsentry_as(w, ts_dt, &t0);
	const void* hd = tsentry_as(ld, ts_dt, &t1);
	const void* hu = tsentry_as(lu, ts_dt, &t2);
	if (!hw || !hd || !hu) { mlis_err_set("lora: unsupported dtype for %s", w->key); R = -1; goto end; }
	size_t bw = (size_t)n0 * n1 * es, bd = (size_t)n0 * r * es, bu = (size_t)n1 * r * es;
	/* the index guarantees nbytes == count x element size (tstore.c), so these reads stay inside the mappings */
	dev = ggml_b200_malloc(bw + bd + bu + 64);
	if (!dev) { mlis_err_set("lora: device allocation failed"); R = -1; goto end; }
	char* ddown = dev + (bw + 15) / 16 * 16; char* dup = ddown + (bd + 15) / 16 * 16;
	ggml_b200_upload(dev, hw, bw); ggml_b200_upload(ddown, hd, bd); ggml_b200_upload(dup, hu, bu);
	if (ts_dt == TS_F32) ggml_b200_lora_merge_f32(dev, ddown, dup, n0, n1, (int)r, scale);
	else ggml_b200_lora_merge_f16(dev, ddown, dup, n0, n1, (int)r, scale);
	merged = xmalloc(bw);
	ggml_b200_download(merged, dev, bw);
	/* NaN/Inf guard on the first element, like lora.c:80-87 */
	float first;
	if (ts_dt == TS_F32) memcpy(&first, merged, 4); else { _Float16 h; memcpy(&h, merged, 2); first = (float)h; }
	if (!(first - first == 0)) { mlis_err_set("NaN in LoRA result"); R = -1; goto end; }
	int64_t shape[4] = { w->shape[0], w->shape[1], w->shape[2], w->shape[3] };
	int nd = w->ndim;
	char* wkey = xstrdup(w->key);
	tstore_add(dst, wkey, ts_dt, nd, shape, merged, bw, true);      /* the store owns `merged` now */
	merged = NULL;
	free(wkey);
end:
	if (dev) ggml_b200_free(dev);
	free(merged); free(t0); free(t1); free(t2);
	return R;
}

int lora_apply(TStore* dst, TStore* lora, float mult, int ts_dt)
{
	static const char suffix[] = ".lora_down.weight";
	char key[320];
	int n_applied = 0;
	for (int i = 0; i < lora->n; ++i) {
		TSEntry* ld = &lora->e[i];
		size_t kl = strlen(ld->key), sl = sizeof(suffix) - 1;
		if (kl <= sl || strcmp(ld->key + kl - sl, suffix)) continue;
		int base = (int)(kl - sl);
		snprintf(key, sizeof(key), "%.*s.weight", base, ld->key);
		TSEntry* w = tstore_find(dst, key);
		if (!w) FAIL(-1, "lora tensor not found in model: %s", key);
		snprintf(key, sizeof(key), "%.*s.lora_up.weight", base, ld->key);
		TSEntry* lu = tstore_find(lora, key);
		if (!lu) FAIL(-1, "lora up tensor not found: %s", key);
		snprintf(key, sizeof(key), "%.*s.scale", base, ld->key);
		TSEntry* ls = tstore_find(lora, key);
		snprintf(key, sizeof(key), "%.*s.alpha", base, ld->key);
		TSEntry* la = tstore_find(lora, key);

		/* shapes in ggml order: down [n0.., r], up [r.., n1] with r the outer dim of down (lora.c:15-26) */
		int64_t r = ld->ndim ? ld->shape[ld->ndim - 1] : 0;
		if (r <= 0) FAIL(-1, "lora down tensor has no rank dimension: %s", ld->key);
		int64_t n0 = tsentry_count(ld) / r, n1 = tsentry_count(lu) / r;
		if (w->ndim < 2 || ld->ndim != w->ndim || lu->ndim != w->ndim || tsentry_count(w) != n0 * n1 || tsentry_count(lu) != n1 * r)
			FAIL(-1, "lora up/down invalid shapes for %s", w->key);
		float scale = 1;
		if (ls) scale = scalar_of(ls);
		else if (la) scale = scalar_of(la) / r;
		scale *= mult;
		if (!(scale > 0)) FAIL(-1, "lora scale must be positive (%g)", scale);
		CHECK(lora_apply_one(dst, w, ld, lu, scale, ts_dt));
		n_applied++;
	}
	return n_applied;
}
