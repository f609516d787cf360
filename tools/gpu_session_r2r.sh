#!/bin/bash
# attn_ap_kernel timeline from the -DATTN_AP_TRACE build (mlimgsynth_b200/build/trace)
TAG=${1:-r2r}
mkdir -p gpurun_out
A=mlimgsynth_b200/build/attn_trace_tr
export GGML_B200_ATTN_SPLIT=5
{
for cfg in "40 4096 4096 8 16" "64 4096 4096 10 4"; do
  for v in "2 1"; do
    set -- $v
    echo "== $cfg split=5 pk=$1 poly=$2"
    GGML_B200_ATTN_PK=$1 GGML_B200_ATTN_POLY=$2 timeout 60 $A $cfg 40 2>&1
  done
done
} > gpurun_out/attn_$TAG.log 2>&1
grep "^==\|us \|max abs" gpurun_out/attn_$TAG.log | paste - - - | awk '{print $2,$3,$4,$5,$6,$7,$8,$9,$10, $17, $18, $NF}'
