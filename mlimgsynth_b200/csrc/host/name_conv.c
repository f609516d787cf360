/* name_conv.c -- checkpoint tensor names -> module-graph parameter paths.
 * Covers the layouts listed in SURVEY.md Appendix B (behaviour of the reference's
 * tensor_name_conv.c:274 `tnconv_sd` for CompVis/LDM UNet+VAE, HF-CLIP and OpenCLIP text encoders).
 * In patterns a '.' also matches '_' or '/' in the input, so LoRA keys
 * ("lora_unet_input_blocks_1_1_transformer_blocks_0_attn1_to_q") go through the same rules
 * (tensor_name_conv.c:6-21). Written as a cursor-based rewriter: each rule consumes a prefix of
 * the input and appends its replacement to the output. */
#include "tstore.h"

typedef struct { const char* in; char* out; size_t cap, len; } Cur;

static bool sep(char c) { return c == '.' || c == '_' || c == '/'; }

static bool peek(const Cur* c, const char* pat)
{
	const char* s = c->in;
	for (; *pat; ++pat, ++s) {
		if (*s == *pat) continue;
		if (*pat == '.' && sep(*s)) continue;
		return false;
	}
	return true;
}
static void put(Cur* c, const char* s)
{
	size_t n = strlen(s);
	if (c->len + n + 1 > c->cap) n = c->cap > c->len + 1 ? c->cap - c->len - 1 : 0;
	memcpy(c->out + c->len, s, n); c->len += n; c->out[c->len] = 0;
}
/* consume `pat`, emit `rep` */
static bool rw(Cur* c, const char* pat, const char* rep)
{
	if (!peek(c, pat)) return false;
	c->in += strlen(pat);
	put(c, rep);
	return true;
}
static bool keep(Cur* c, const char* pat) { return rw(c, pat, pat); }
/* consume "<digits><sep>", emit "<digits>." ; optionally return the number */
static bool num(Cur* c, int* value, bool emit)
{
	const char* s = c->in;
	if (*s < '0' || *s > '9') return false;
	int v = 0;
	while (*s >= '0' && *s <= '9') v = v * 10 + (*s++ - '0');
	if (!sep(*s)) return false;
	if (emit) { char b[16]; snprintf(b, sizeof(b), "%d.", v); put(c, b); }
	if (value) *value = v;
	c->in = s + 1;
	return true;
}
static int tail(Cur* c, int r) { if (r > 0) put(c, c->in); return r; }

static int clip_hf(Cur* c)      /* transformer.text_model.* */
{
	if (!rw(c, "transformer.text_model.", "text.")) return 0;
	if (rw(c, "embeddings.", "embed.")) {
		if (rw(c, "position_embedding.", "position.") || rw(c, "token_embedding.", "token.")) return tail(c, 1);
		return 0;
	}
	if (keep(c, "encoder.layers.")) {
		num(c, NULL, true);
		if (rw(c, "layer_norm1.", "norm1.") || rw(c, "layer_norm2.", "norm2.") || rw(c, "self_attn.", "attn.") || keep(c, "mlp."))
			return tail(c, 1);
		return 0;
	}
	if (rw(c, "final_layer_norm.", "ln_final.") || rw(c, "text_projection", "text_proj")) return tail(c, 1);
	return 0;
}

static int clip_open(Cur* c)    /* model.* (OpenCLIP) */
{
	if (!rw(c, "model.", "text.")) return 0;
	if (keep(c, "ln_final.") || rw(c, "token_embedding.", "embed.token.") ||
		rw(c, "positional_embedding", "embed.position.weight") || rw(c, "text_projection", "text_proj"))
		return tail(c, 1);
	if (rw(c, "transformer.resblocks.", "encoder.layers.")) {
		num(c, NULL, true);
		if (rw(c, "ln_1.", "norm1.") || rw(c, "ln_2.", "norm2.") || rw(c, "mlp.c_fc.", "mlp.fc1.") || rw(c, "mlp.c_proj.", "mlp.fc2."))
			return tail(c, 1);
		if (keep(c, "attn.")) {
			if (peek(c, "in_proj_bias") || peek(c, "in_proj_weight")) return tail(c, 2);
			if (keep(c, "out_proj.")) return tail(c, 1);
		}
	}
	return 0;
}

static int vae(Cur* c)
{
	bool dec = keep(c, "decoder."), enc = !dec && keep(c, "encoder.");
	if (dec || enc) {
		Cur save = *c;
		if (keep(c, dec ? "up." : "down.") && num(c, NULL, true) && keep(c, "block.") && num(c, NULL, true))
			rw(c, "nin_shortcut.", "skip_conv.");
		else { size_t l = save.len; *c = save; c->out[l] = 0; }
		return tail(c, 1);
	}
	if (keep(c, "quant_conv.") || keep(c, "post_quant_conv.")) return tail(c, 1);
	return 0;
}

static int unet_block(Cur* c)
{
	if (rw(c, "transformer_blocks.", "transf.")) {
		num(c, NULL, true);
		if (keep(c, "attn1.") || keep(c, "attn2.")) {
			(void)(rw(c, "to_q.", "q_proj.") || rw(c, "to_k.", "k_proj.") || rw(c, "to_v.", "v_proj.") || rw(c, "to_out.0.", "out_proj."));
			return tail(c, 1);
		}
		if (keep(c, "ff.")) { if (keep(c, "net.0.") || keep(c, "net.2.")) return tail(c, 1); return 0; }
		if (keep(c, "norm1.") || keep(c, "norm2.") || keep(c, "norm3.")) return tail(c, 1);
		return 0;
	}
	if (rw(c, "in_layers.0.", "norm1.") || rw(c, "in_layers.2.", "conv1.") || rw(c, "out_layers.0.", "norm2.") ||
		rw(c, "out_layers.3.", "conv2.") || rw(c, "emb_layers.1.", "emb_proj.") || rw(c, "skip_connection.", "skip_conv.") ||
		rw(c, "op.", "conv.") || keep(c, "norm.") || keep(c, "proj_in.") || keep(c, "proj_out.") || keep(c, "conv."))
		return tail(c, 1);
	return 0;
}

static int unet(Cur* c)
{
	if (keep(c, "time_embed.") || rw(c, "label_emb.0.", "label_embed.") || rw(c, "input_blocks.0.0.", "in.conv.") ||
		rw(c, "out.0.", "out.norm.") || rw(c, "out.2.", "out.conv."))
		return tail(c, 1);
	if ((rw(c, "input_blocks.", "in.") && num(c, NULL, true)) || (rw(c, "output_blocks.", "out.") && num(c, NULL, true)) ||
		rw(c, "middle_block.", "mid.")) {
		num(c, NULL, true);
		return unet_block(c);
	}
	return 0;
}

int tnconv_sd(const char* key, char* out, size_t out_sz)
{
	Cur c = { key, out, out_sz, 0 };
	out[0] = 0;
	if (rw(&c, "cond_stage_model.", "clip.")) {
		if (peek(&c, "transformer.text_model.")) return clip_hf(&c);
		if (peek(&c, "model.")) return clip_open(&c);
		return 0;
	}
	if (rw(&c, "conditioner.embedders.0.", "clip.")) return clip_hf(&c);
	if (rw(&c, "conditioner.embedders.1.", "clip2.")) return clip_open(&c);
	if (rw(&c, "first_stage_model.", "vae.")) return vae(&c);
	if (rw(&c, "model.diffusion_model.", "unet.") || keep(&c, "unet.")) return unet(&c);
	return 0;
}
