#!/bin/bash
TAG=${1:-r2d}
mkdir -p gpurun_out
A=mlimgsynth_b200/build/attn_trace
GGML_B200_ATTN_SPLIT=1 timeout 120 $A 40 4096 4096 8 16 8 > gpurun_out/attn_trace_split_$TAG.log 2>&1; cat gpurun_out/attn_trace_split_$TAG.log
GGML_B200_ATTN_SPLIT=1 timeout 120 $A 64 4096 4096 10 4 6 > gpurun_out/attn_trace_split64_$TAG.log 2>&1; cat gpurun_out/attn_trace_split64_$TAG.log
