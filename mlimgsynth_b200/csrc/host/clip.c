#define _GNU_SOURCE
#include "clip.h"
#include "mlblock_nn.h"
#include "unicode_tables.h"
#include <dlfcn.h>

#define N(name, x) mlctx_tensor_add(C, (name), (x))

const ClipParams g_clip_vit_l_14    = { 49408, 77,  768, 3072, 12, 12, 49406, 49407, 49407 };
const ClipParams g_clip_vit_h_14    = { 49408, 77, 1024, 4096, 16, 24, 49406, 49407, 0 };
const ClipParams g_clip_vit_bigg_14 = { 49408, 77, 1280, 5120, 20, 32, 49406, 49407, 0 };

/* ------------------------------------------------------------------ tokenizer
 * OpenAI CLIP simple_tokenizer semantics as implemented by the reference:
 *  - tokens 0..255 are the byte alphabet, 256..511 the same with the end-of-word mark,
 *    512 + r is the result of merge r (clip.c:80-141);
 *  - text is cut into words: a run of letters, a run of digits (the reference groups digit runs,
 *    clip.c:213-253), or a run of anything else; ASCII white space and Unicode separators split;
 *    the contractions 's 't 're 've 'm 'll (matched case-insensitively) are words of their own -- 'd is
 *    NOT in the list (clip.c:228-233);
 *  - each word is lower-cased per code point, turned into byte tokens, the last one marked
 *    end-of-word, then greedily merged by lowest merge rank (first position wins ties). */
#define N_MERGES 48894
static uint16_t (*g_merges)[2];          /* [N_MERGES][2] */
static uint32_t* g_mkey; static int32_t* g_mval; static uint32_t g_mcap;   /* pair -> rank hash */

static uint32_t pair_hash(uint32_t k) { k ^= k >> 15; k *= 0x2c1b3c6du; k ^= k >> 12; k *= 0x297a2d39u; k ^= k >> 15; return k; }

int clip_tokenizer_load(const char* dir)
{
	if (g_merges) return 1;
	/* candidates in order: option aux_dir (the reference CLI points it at the directory of its binary,
	   main_mlimgsynth.c:642-651), $MLIS_B200_DATA, then <directory of this library>/../data */
	char path[1024], tried[3200]; tried[0] = 0;
	const char* env = getenv("MLIS_B200_DATA");
	FILE* f = NULL;
	for (int c = 0; c < 3 && !f; ++c) {
		path[0] = 0;
		if (c == 0 && dir && *dir) snprintf(path, sizeof(path), "%s/clip_merges.bin", dir);
		else if (c == 1 && env && *env) snprintf(path, sizeof(path), "%s/clip_merges.bin", env);
		else if (c == 2) {
			Dl_info info;
			if (dladdr((void*)clip_tokenizer_load, &info) && info.dli_fname) {
				snprintf(path, sizeof(path), "%s", info.dli_fname);
				char* s = strrchr(path, '/');
				if (s) *s = 0; else strcpy(path, ".");
				strncat(path, "/../data/clip_merges.bin", sizeof(path) - strlen(path) - 1);
			}
		}
		if (!path[0]) continue;
		f = fopen(path, "rb");
		if (!f) { strncat(tried, path, sizeof(tried) - strlen(tried) - 2); strncat(tried, " ", sizeof(tried) - strlen(tried) - 1); }
	}
	if (!f) FAIL(-6, "CLIP merge table not found (tried: %s; set option aux_dir or MLIS_B200_DATA)", tried);
	g_merges = xmalloc(N_MERGES * 4);
	size_t got = fread(g_merges, 4, N_MERGES, f);
	fclose(f);
	if (got != N_MERGES) { free(g_merges); g_merges = NULL; FAIL(-6, "%s: truncated merge table", path); }
	g_mcap = 1u << 17;
	g_mkey = xmalloc(g_mcap * 4); g_mval = xmalloc(g_mcap * 4);
	memset(g_mval, 0xff, g_mcap * 4);
	for (int r = 0; r < N_MERGES; ++r) {
		uint32_t k = ((uint32_t)g_merges[r][0] << 16) | g_merges[r][1], h = pair_hash(k) & (g_mcap - 1);
		while (g_mval[h] >= 0) h = (h + 1) & (g_mcap - 1);
		g_mkey[h] = k; g_mval[h] = r;
	}
	return 1;
}

static int32_t merge_rank(int32_t left, int32_t right)
{
	if (left > 0xffff || right > 0xffff) return INT32_MAX;
	uint32_t k = ((uint32_t)left << 16) | (uint32_t)right, h = pair_hash(k) & (g_mcap - 1);
	while (g_mval[h] >= 0) { if (g_mkey[h] == k) return g_mval[h]; h = (h + 1) & (g_mcap - 1); }
	return INT32_MAX;
}

/* byte <-> alphabet token (the printable-first ordering of CLIP's bytes_to_unicode, clip.c:117-141) */
static int byte_to_token(uint8_t b)
{
	if (b <= 32) return b + 188;
	if (b <= 126) return b - 33;
	if (b <= 160) return b + 94;
	if (b <= 172) return b - 67;
	if (b == 173) return 255;
	return b - 68;
}
static int token_to_byte(int t)
{
	if (t <= 93) return t + 33;
	if (t <= 105) return t + 67;
	if (t <= 187) return t + 68;
	if (t <= 220) return t - 188;
	if (t <= 254) return t - 94;
	if (t == 255) return 173;
	return -1;
}

static uint32_t utf8_next(const char** p, const char* end)
{
	const uint8_t* s = (const uint8_t*)*p;
	uint32_t c = *s++;
	int extra = c >= 0xF0 ? 3 : c >= 0xE0 ? 2 : c >= 0xC0 ? 1 : 0;
	if (extra) {
		c &= 0x3F >> extra;
		while (extra-- && (const char*)s < end && (*s & 0xC0) == 0x80) c = (c << 6) | (*s++ & 0x3F);
	}
	*p = (const char*)s;
	return c;
}
static int utf8_put(char* out, uint32_t c)
{
	if (c < 0x80) { out[0] = (char)c; return 1; }
	if (c < 0x800) { out[0] = (char)(0xC0 | (c >> 6)); out[1] = (char)(0x80 | (c & 0x3F)); return 2; }
	if (c < 0x10000) { out[0] = (char)(0xE0 | (c >> 12)); out[1] = (char)(0x80 | ((c >> 6) & 0x3F)); out[2] = (char)(0x80 | (c & 0x3F)); return 3; }
	out[0] = (char)(0xF0 | (c >> 18)); out[1] = (char)(0x80 | ((c >> 12) & 0x3F)); out[2] = (char)(0x80 | ((c >> 6) & 0x3F)); out[3] = (char)(0x80 | (c & 0x3F));
	return 4;
}
static bool in_ranges(const uint32_t (*tab)[2], int n, uint32_t cp)
{
	int lo = 0, hi = n;
	while (lo < hi) { int m = (lo + hi) / 2; if (tab[m][1] < cp) lo = m + 1; else hi = m; }
	return lo < n && tab[lo][0] <= cp;
}
#define COUNT(a) ((int)(sizeof(a) / sizeof((a)[0])))
static uint32_t uc_to_lower(uint32_t cp)
{
	int lo = 0, hi = COUNT(uc_lower);
	while (lo < hi) { int m = (lo + hi) / 2; if (uc_lower[m][0] < cp) lo = m + 1; else hi = m; }
	return (lo < COUNT(uc_lower) && uc_lower[lo][0] == cp) ? uc_lower[lo][1] : cp;
}
static bool ascii_space(uint32_t c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\v' || c == '\f'; }
/* character class: 'Z' separator, 'L' letter, 'N' number, 'P' everything else */
static int char_class(uint32_t cp)
{
	if (ascii_space(cp) || in_ranges(uc_Z, COUNT(uc_Z), cp)) return 'Z';
	if (in_ranges(uc_L, COUNT(uc_L), cp)) return 'L';
	if (in_ranges(uc_N, COUNT(uc_N), cp)) return 'N';
	return 'P';
}

static int match_contraction(const char* cur, const char* end)
{
	static const char* list[] = { "'s", "'t", "'re", "'ve", "'m", "'ll", NULL };
	for (int i = 0; list[i]; ++i) {
		const char *s = list[i], *c = cur;
		for (; c < end && *s; ++c, ++s) { int ch = (*c >= 'A' && *c <= 'Z') ? *c + 32 : *c; if (ch != *s) break; }
		if (!*s) return (int)(c - cur);
	}
	return 0;
}

/* next word of [*pcur, end): returns its [beg, *pcur) span */
static const char* next_word(const char** pcur, const char* end)
{
	const char* cur = *pcur;
	while (cur < end) {   /* skip separators */
		const char* p = cur;
		if (char_class(utf8_next(&p, end)) != 'Z') break;
		cur = p;
	}
	const char* beg = cur;
	int cls = 0;
	while (cur < end) {
		int m = match_contraction(cur, end);
		if (m) { if (!cls) cur += m; break; }
		const char* p = cur;
		int c = char_class(utf8_next(&p, end));
		if (c == 'Z' || (cls && c != cls)) break;
		cls = c; cur = p;
	}
	*pcur = cur;
	return beg;
}

static int bpe_word(const char* w, const char* wend, int32_t* tok, int max)
{
	int n = 0;
	char buf[4];
	while (w < wend) {
		uint32_t cp = uc_to_lower(utf8_next(&w, wend));
		int nb = utf8_put(buf, cp);
		for (int i = 0; i < nb; ++i) { if (n == max) FAIL(-1, "word too long"); tok[n++] = byte_to_token((uint8_t)buf[i]); }
	}
	if (!n) return 0;
	tok[n - 1] += 256;
	while (n > 1) {
		int32_t best = INT32_MAX; int pos = 0;
		for (int i = 1; i < n; ++i) { int32_t r = merge_rank(tok[i-1], tok[i]); if (r < best) { best = r; pos = i; } }
		if (best == INT32_MAX) break;
		tok[pos - 1] = 512 + best;
		memmove(tok + pos, tok + pos + 1, (n - pos - 1) * sizeof(*tok));
		n--;
	}
	return n;
}

int clip_tokenize(const ClipParams* P, const char* text, size_t len, int32_t** ptok, int* pn, int* pcap)
{
	(void)P;
	CHECK(clip_tokenizer_load(NULL));
	const char *cur = text, *end = text + len;
	int need = *pn + (int)len + 1;
	if (need > *pcap) { *pcap = need; *ptok = xrealloc(*ptok, (size_t)need * sizeof(int32_t)); }
	for (;;) {
		const char* beg = next_word(&cur, end);
		if (cur == beg) break;
		int r = bpe_word(beg, cur, *ptok + *pn, *pcap - *pn);
		if (r < 0) return r;
		*pn += r;
	}
	return *pn;
}

int clip_token_decode(const ClipParams* P, int32_t token, size_t bufsz, char* buf)
{
	if (token < 0 || !g_merges) return -1;
	if (token < 256) { if (bufsz < 1) return -1; buf[0] = (char)token_to_byte(token); return 1; }
	if (token < 512) { if (bufsz < 2) return -1; buf[0] = (char)token_to_byte(token - 256); buf[1] = ' '; return 2; }
	if (token >= 512 + N_MERGES) return -1;
	int a = clip_token_decode(P, g_merges[token - 512][0], bufsz, buf);
	if (a < 0) return a;
	int b = clip_token_decode(P, g_merges[token - 512][1], bufsz - a, buf + a);
	return b < 0 ? b : a + b;
}

/* ------------------------------------------------------------------ text transformer */
static MLTensor* clip_layer(MLCtx* C, MLTensor* x, const ClipParams* P)
{
	mlctx_block_begin(C);
	MLTensor* h = N("norm1", mlb_nn_layer_norm(C, x, true, true, 0));
	x = ggml_add(C->cc, x, N("attn", mlb_attn_mhead(C, h, h, h, P->d_embed, P->d_embed, P->n_head, true, true, true)));
	h = N("norm2", mlb_nn_layer_norm(C, x, true, true, 0));
	mlctx_block_begin(C);   /* mlp */
	h = N("fc1", mlb_nn_linear(C, h, P->n_interm, true));
	/* OpenCLIP towers (d 1024 / 1280) use GELU, the OpenAI L/14 tower quick-GELU (clip.c:353-357) */
	h = (P->d_embed == 1024 || P->d_embed == 1280) ? ggml_gelu_inplace(C->cc, h) : ggml_gelu_quick_inplace(C->cc, h);
	h = N("mlp", N("fc2", mlb_nn_linear(C, h, P->d_embed, true)));
	return ggml_add(C->cc, x, h);
}

MLTensor* mlb_clip_text(MLCtx* C, MLTensor* tokens, const ClipParams* P, int clip_skip, bool norm)
{
	char name[32];
	mlctx_block_begin(C);
	/* embeddings (clip.c:319-344): token rows + learned positions */
	mlctx_block_begin(C);
	MLTensor* tw = N("token.weight", ggml_new_tensor_2d(C->cp, C->c.wtype, P->d_embed, P->n_vocab));
	MLTensor* pw = N("position.weight", ggml_new_tensor_2d(C->cp, GGML_TYPE_F32, P->d_embed, P->n_token));
	MLTensor* x = ggml_reshape_3d(C->cc, tokens, tokens->ne[0], 1, tokens->ne[1]);
	x = ggml_get_rows(C->cc, tw, x);
	x = ggml_reshape_3d(C->cc, x, x->ne[0], x->ne[1], x->ne[3]);
	x = N("embed", ggml_add(C->cc, x, pw));
	/* encoder: pre-LN causal layers; clip_skip drops the last (clip_skip-1) layers */
	int n_layer = P->n_layer - (clip_skip > 1 ? clip_skip - 1 : 0);
	mlctx_block_begin(C);
	for (int i = 0; i < n_layer; ++i) {
		snprintf(name, sizeof(name), "layers.%d", i);
		x = N(name, clip_layer(C, x, P));
	}
	x = N("encoder", x);
	if (norm) x = N("ln_final", mlb_nn_layer_norm(C, x, true, true, 0));
	return x;
}

MLTensor* mlb_clip_text_proj(MLCtx* C, MLTensor* x, int i_tok_end)
{
	int d = (int)x->ne[0];
	MLTensor* p = N("text_proj", ggml_new_tensor_2d(C->cp, GGML_TYPE_F32, d, d));
	p = ggml_cont(C->cc, ggml_transpose(C->cc, p));
	/* i_tok_end >= 0: features of the first end-of-text token only (clip.c:418-437). i_tok_end < 0: project every
	   position; the caller picks the row it needs, so the graph does not depend on the prompt length. */
	if (i_tok_end >= 0) x = ggml_view_1d(C->cc, x, d, x->nb[1] * i_tok_end);
	return ggml_mul_mat(C->cc, p, x);
}

int clip_text_encode(ClipState* S, MLCtx* C, const ClipParams* P, const char* tprefix, unsigned n_tok, const int32_t* toks,
	HTensor* embed, HTensor* feat, int clip_skip, bool norm)
{
	if (feat) { clip_skip = -1; norm = true; }
	if (n_tok + 2 > (unsigned)P->n_token) FAIL(-1, "prompt too long (max: %d)", P->n_token - 2);
	int32_t tokens[128];
	tokens[0] = P->tok_start;
	memcpy(tokens + 1, toks, n_tok * sizeof(int32_t));
	tokens[n_tok + 1] = P->tok_end;
	for (int i = n_tok + 2; i < P->n_token; ++i) tokens[i] = P->tok_pad;

	/* The reference's pooled-feature graph depends on the EOS position (view at row n_tok + 1, clip.c:431), which would
	   rebuild the graph -- and re-upload the encoder's weights -- whenever the prompt length changes (every generation:
	   prompt and negative prompt alternate). Here every position is projected (row-wise independent, same arithmetic
	   per row) and the EOS row is read back, so one graph serves all prompts. */
	bool reuse = C->prepared && S->ctx == C && S->par == P && S->clip_skip == clip_skip && S->norm == norm &&
		S->with_feat == (feat != NULL);
	if (!reuse) {
		mlctx_begin(C, "CLIP text encode");
		S->t_tok = mlctx_input_new(C, "tokens", GGML_TYPE_I32, P->n_token, 1, 1, 1);
		S->t_embed = mlb_clip_text(C, S->t_tok, P, clip_skip, norm);
		MLTensor* result = S->t_embed;
		S->t_feat = NULL;
		/* both results come from the same graph when features are requested, as in the reference (clip.c:439-488) */
		ggml_set_output(S->t_embed);
		if (feat) result = S->t_feat = mlb_clip_text_proj(C, S->t_embed, -1);
		mlctx_tensor_add(C, "text", result);
		C->c.tprefix = tprefix;
		CHECK(mlctx_prep(C));
		S->ctx = C; S->par = P; S->clip_skip = clip_skip; S->norm = norm; S->with_feat = feat != NULL; S->n_tok_feat = n_tok;
	}
	ggml_backend_tensor_set(S->t_tok, tokens, 0, sizeof(int32_t) * P->n_token);
	CHECK(mlctx_compute(C));
	if (embed) {
		ht_resize(embed, P->d_embed, P->n_token, 1, 1);
		ggml_backend_tensor_get(S->t_embed, embed->d, 0, ht_count(embed) * sizeof(float));
	}
	if (feat) {
		ht_resize(feat, P->d_embed, 1, 1, 1);
		ggml_backend_tensor_get(S->t_feat, feat->d, (size_t)(n_tok + 1) * P->d_embed * sizeof(float), ht_count(feat) * sizeof(float));
	}
	return 1;
}
