/* lora.h -- LoRA merge, applied on the device (reference: lora.c:9-138). */
#pragma once
#include "tstore.h"
/* For every "<key>.lora_down.weight" in ts_lora: W(<key>.weight) += scale * up . down with
 * scale = (".scale" | ".alpha"/rank) * mult; the merged weight replaces the entry in ts_dst. */
/* ts_dtype: TS_F16 or TS_F32 = the context's weight type (the reference merges in C->c.wtype, lora.c:46-50). */
int lora_apply(TStore* ts_dst, TStore* ts_lora, float mult, int ts_dtype);
/* name conversion for LoRA files: "lora_unet_input_blocks_1_1_..._to_q.lora_down.weight" ->
 * "unet.in.1.1....q_proj.lora_down.weight" (mlimgsynth.c:1067-1092) */
int lora_name_conv(const char* file_key, char* out, size_t out_sz);
