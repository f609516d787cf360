#!/bin/bash
TAG=${1:-r2h}
mkdir -p gpurun_out
timeout 600 python tools/phase_times.py 8 2>&1 | tee gpurun_out/phase_times_$TAG.log
GGML_B200_ATTN_SPLIT=2 timeout 120 mlimgsynth_b200/build/attn_trace 40 4096 4096 8 16 2 2>&1 | grep -v "^blk\|^softmax\|^mma" | tee gpurun_out/attn_clock_$TAG.log
GGML_B200_ATTN_SPLIT=2 timeout 120 mlimgsynth_b200/build/attn_trace 64 4096 4096 10 4 2 2>&1 | grep -v "^blk\|^softmax\|^mma" | tee -a gpurun_out/attn_clock_$TAG.log
