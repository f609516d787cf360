#!/bin/bash
# packed-softmax variants of the key-halves attention kernel (GGML_B200_ATTN_PK) on the three benchmark shapes + edge shapes
TAG=${1:-r2n}
mkdir -p gpurun_out
A=mlimgsynth_b200/build/attn_trace
{
for cfg in "40 4096 4096 8 16" "64 4096 4096 10 4" "64 9216 9216 5 2" "40 4096 4000 8 2" "64 1000 1090 3 2" "48 300 200 2 1"; do
  for v in "0 1" "1 0" "1 1" "1 2" "1 3" "2 0" "2 1" "2 2" "2 3" "2 4"; do
    set -- $v
    echo "== $cfg pk=$1 poly=$2"
    GGML_B200_ATTN_PK=$1 GGML_B200_ATTN_POLY=$2 timeout 120 $A $cfg 0 2>&1 | grep -v "^softmax\|^blk\|^mma\|^clock"
  done
done
} > gpurun_out/attn_$TAG.log 2>&1
grep -A2 "^==" gpurun_out/attn_$TAG.log | grep -v "^--" | paste - - - | awk '{print $2,$3,$4,$5,$6,$7,$8, $15, $16, $NF}'
