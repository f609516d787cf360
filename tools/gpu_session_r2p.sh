#!/bin/bash
# attn_ap_kernel timeline (CTA 0) + correctness after the epilogue wait fix
TAG=${1:-r2p}
mkdir -p gpurun_out
A=mlimgsynth_b200/build/attn_trace
export GGML_B200_ATTN_SPLIT=5
{
for cfg in "40 4096 4096 8 16" "64 4096 4096 10 4"; do
  for v in "2 2 0" "2 2 1"; do
    set -- $v
    echo "== $cfg split=5 pk=$1 poly=$2 sig=$3"
    GGML_B200_ATTN_PK=$1 GGML_B200_ATTN_POLY=$2 GGML_B200_ATTN_SIG=$3 timeout 60 $A $cfg 30 2>&1
  done
done
for cfg in "40 1024 1024 8 16" "64 9216 9216 5 2" "40 4096 4000 8 2" "64 1000 1090 3 2" "48 300 200 2 1" "40 300 4096 8 2" "64 128 256 1 1" "64 129 257 1 1" "64 1024 191 2 1" "40 1024 193 2 1"; do
  for v in "2 2 0" "1 3 1"; do
    set -- $v
    echo "== $cfg split=5 pk=$1 poly=$2 sig=$3"
    GGML_B200_ATTN_PK=$1 GGML_B200_ATTN_POLY=$2 GGML_B200_ATTN_SIG=$3 timeout 60 $A $cfg 0 2>&1 | grep -v "^clock"
  done
done
} > gpurun_out/attn_$TAG.log 2>&1
grep "^==\|us \|max abs" gpurun_out/attn_$TAG.log | paste - - - | awk '{print $2,$3,$4,$5,$6,$7,$8,$9,$10, $17, $18, $NF}'
