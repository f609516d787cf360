#!/bin/bash
TAG=${1:-r2j}
mkdir -p gpurun_out
A=mlimgsynth_b200/build/attn_trace
{
for cfg in "40 4096 4096 8 16" "64 4096 4096 10 4" "64 9216 9216 5 2" "40 4096 4000 8 2" "64 1000 1090 3 2" "48 300 200 2 1"; do
  for v in "3 1 1" "3 0 1" "3 1 0" "4 0 0" "2 1 1"; do
    set -- $v
    echo "== $cfg split=$1 poly=$2 park=$3"
    GGML_B200_ATTN_SPLIT=$1 GGML_B200_ATTN_POLY=$2 GGML_B200_ATTN_PARK=$3 timeout 120 $A $cfg 0 2>&1 | grep -v "^softmax\|^blk\|^mma"
  done
done
} > gpurun_out/attn_$TAG.log 2>&1
grep -A1 "^==" gpurun_out/attn_$TAG.log | grep -v "^--" | paste - - | awk '{print $2,$3,$4,$5,$6,$7,$8,$9, $15, $16}'
grep "max abs err" gpurun_out/attn_$TAG.log | sort | uniq -c | sort -k6 -g | tail -3
