/* prompt_preproc.h -- A1111-style prompt parser: emphasis "()" "[]" "(text:1.5)", "<lora:NAME:MULT>",
 * backslash escapes, BREAK. Bit-exact behaviour of the reference's prompt_preproc.h:105-209
 * (weights are pow(1.1, n_paren - n_bracket) evaluated in double and stored as float). */
#pragma once
#include "base.h"

typedef struct PromptChunk { int beg, len; float w; } PromptChunk;        /* span of PromptText.text */
typedef struct PromptLora  { int beg, len; float w; } PromptLora;         /* span of PromptText.data */
typedef struct PromptText {
	char* text; int text_len, text_cap;
	char* data; int data_len, data_cap;
	PromptChunk* chunks; int n_chunks, cap_chunks;
	PromptLora* loras; int n_loras, cap_loras;
} PromptText;

void prompt_text_free(PromptText* S);
void prompt_text_clear(PromptText* S);
void prompt_text_set_raw(PromptText* S, const char* s, size_t len);
int  prompt_text_set_parse(PromptText* S, const char* s, size_t len);   /* < 0: MLIS_E_PROMPT_PARSE (-5) */
