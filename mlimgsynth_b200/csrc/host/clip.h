/* clip.h -- CLIP text conditioning: byte-level BPE tokenizer (bit-exact with the reference's
 * clip.c:59-315, including its quirks) and the text transformer graph (clip.c:319-437). */
#pragma once
#include "mlblock.h"

typedef struct ClipParams {
	int n_vocab, n_token, d_embed, n_interm, n_head, n_layer;
	uint32_t tok_start, tok_end, tok_pad;
} ClipParams;

extern const ClipParams g_clip_vit_l_14, g_clip_vit_h_14, g_clip_vit_bigg_14;

/* Locate and load the merge table (data/clip_merges.bin); dir may be NULL (search next to the library). */
int clip_tokenizer_load(const char* dir);
/* Appends the token ids of `text` (no BOS/EOS) to *ptok (malloc'd, *pn used, *pcap capacity). */
int clip_tokenize(const ClipParams* P, const char* text, size_t len, int32_t** ptok, int* pn, int* pcap);
int clip_token_decode(const ClipParams* P, int32_t token, size_t bufsz, char* buf);

MLTensor* mlb_clip_text(MLCtx* C, MLTensor* tokens, const ClipParams* P, int clip_skip, bool norm);
MLTensor* mlb_clip_text_proj(MLCtx* C, MLTensor* embed, int i_tok_end);

typedef struct ClipState { MLCtx* ctx; const ClipParams* par; int clip_skip; bool norm, with_feat; MLTensor *t_tok, *t_embed, *t_feat; int n_tok_feat; } ClipState;
/* Encode n_tok ids (BOS/EOS/pad added here): embed [d_embed, 77] and/or pooled feat [d_embed].
 * The prepared graph is cached in S and reused while (params, clip_skip, norm) stay the same. */
int clip_text_encode(ClipState* S, MLCtx* C, const ClipParams* P, const char* tprefix, unsigned n_tok, const int32_t* toks,
	HTensor* embed, HTensor* feat, int clip_skip, bool norm);
