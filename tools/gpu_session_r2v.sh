#!/bin/bash
# attn_ap_kernel: which resource bounds it? (GGML_B200_ATTN_DBG: 1 no PV products, 2 no sum products, 4 no QK products, 8 no exponentials)
TAG=${1:-r2v}
mkdir -p gpurun_out
A=mlimgsynth_b200/build/attn_trace
{
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks_throttle_reasons.active,power.draw,temperature.gpu --format=csv
for cfg in "40 4096 4096 8 16" "64 4096 4096 10 4"; do
  echo "== $cfg base"
  GGML_B200_ATTN_PK=2 GGML_B200_ATTN_POLY=2 timeout 30 $A $cfg 0 2>&1 | grep "us "
  for dbg in 0 1 4 7 8 15; do
    echo "== $cfg dbg=$dbg"
    GGML_B200_ATTN_SPLIT=5 GGML_B200_ATTN_PK=2 GGML_B200_ATTN_POLY=1 GGML_B200_ATTN_DBG=$dbg timeout 30 $A $cfg 0 2>&1 | grep "us "
  done
done
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks_throttle_reasons.active,power.draw,temperature.gpu --format=csv
} > gpurun_out/attn_$TAG.log 2>&1
cat gpurun_out/attn_$TAG.log | paste - - | cut -c1-160
