#include "prompt_preproc.h"
#include <math.h>

#define E_PARSE (-5)

void prompt_text_free(PromptText* S) { free(S->text); free(S->data); free(S->chunks); free(S->loras); memset(S, 0, sizeof(*S)); }
void prompt_text_clear(PromptText* S) { S->text_len = S->data_len = S->n_chunks = S->n_loras = 0; if (S->text) S->text[0] = 0; }

static void text_push(PromptText* S, char c)
{
	if (S->text_len + 2 > S->text_cap) { S->text_cap = S->text_cap ? S->text_cap * 2 : 64; S->text = xrealloc(S->text, S->text_cap); }
	S->text[S->text_len++] = c; S->text[S->text_len] = 0;
}
static void chunk_push(PromptText* S, int beg, float w)
{
	PromptChunk c = { beg, 0, w };
	ARR_PUSH(S->chunks, S->n_chunks, S->cap_chunks, c);
}

void prompt_text_set_raw(PromptText* S, const char* s, size_t len)
{
	prompt_text_clear(S);
	for (size_t i = 0; i < len; ++i) text_push(S, s[i]);
	chunk_push(S, 0, 1.0f);
	S->chunks[0].len = S->text_len;
}

/* "<lora:NAME>" or "<lora:NAME:MULT>" */
static int option_parse(PromptText* S, const char* b, const char* e)
{
	if (e - b < 5 || memcmp(b, "lora:", 5)) FAIL(E_PARSE, "prompt: unknown option '%.*s'", (int)(e - b), b);
	b += 5;
	const char* sep = b;
	while (sep < e && *sep != ':') sep++;
	float mult = 1;
	if (sep < e) {
		char* tail = NULL;
		mult = strtof(sep + 1, &tail);
		if (tail != e) FAIL(E_PARSE, "prompt: invalid lora multiplier");
	}
	int len = (int)(sep - b);
	if (S->data_len + len + 1 > S->data_cap) { S->data_cap = (S->data_len + len + 1) * 2; S->data = xrealloc(S->data, S->data_cap); }
	memcpy(S->data + S->data_len, b, len);
	PromptLora l = { S->data_len, len, mult };
	S->data_len += len;
	ARR_PUSH(S->loras, S->n_loras, S->cap_loras, l);
	return 1;
}

int prompt_text_set_parse(PromptText* S, const char* s, size_t len)
{
	prompt_text_clear(S);
	chunk_push(S, 0, 1.0f);
	int n_paren = 0, n_bracket = 0;
	const char* end = s + len;
	for (const char* cur = s; cur < end; ++cur) {
		char c = *cur;
		if (c == '\\') {                       /* escape: next char literally ("\n" -> newline); a trailing '\' is dropped */
			if (cur + 1 < end) { cur++; text_push(S, *cur == 'n' ? '\n' : *cur); }
		}
		else if (c == '(' || c == ')' || c == '[' || c == ']') {
			n_paren += (c == '(') - (c == ')');
			n_bracket += (c == '[') - (c == ']');
			if (n_paren < 0 || n_bracket < 0) FAIL(E_PARSE, "prompt: unmatched ')' or ']'");
			float w = pow(1.1, n_paren - n_bracket);
			PromptChunk* last = &S->chunks[S->n_chunks - 1];
			if (last->beg == S->text_len) last->w = w;           /* empty chunk: just re-weight it */
			else { last->len = S->text_len - last->beg; chunk_push(S, S->text_len, w); }
		}
		else if (c == ':' && (n_paren > 0 || n_bracket > 0)) {   /* explicit weight: only as "(text:w)" at depth 1 */
			if (!(n_paren == 1 && n_bracket == 0)) FAIL(E_PARSE, "prompt: custom emphasis multiplier outside of '()'");
			char* tail = NULL; float w = 0;
			if (cur + 1 < end) { cur++; w = strtof(cur, &tail); }
			if (!(tail && tail < end && *tail == ')')) FAIL(E_PARSE, "prompt: invalid emphasis with ':'");
			cur = tail - 1;
			S->chunks[S->n_chunks - 1].w = w;
		}
		else if (c == '<') {
			const char* e = cur + 1;
			while (e < end && *e != '>') ++e;
			if (e >= end) FAIL(E_PARSE, "prompt: '<' not matched with '>'");
			CHECK(option_parse(S, cur + 1, e));
			cur = e;
		}
		else if (c == 'B' && cur + 5 < end && !memcmp(cur, "BREAK", 5)) cur += 4;   /* dropped (only when text follows) */
		else text_push(S, c);
	}
	PromptChunk* last = &S->chunks[S->n_chunks - 1];
	last->len = S->text_len - last->beg;
	return 1;
}
