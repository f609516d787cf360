#!/bin/bash
# attn_ap_kernel, software-pipelined softmax loop: row-sum mode x polynomial share; correctness on edge shapes
TAG=${1:-r2s}
mkdir -p gpurun_out
A=mlimgsynth_b200/build/attn_trace
{
for cfg in "40 4096 4096 8 16" "64 4096 4096 10 4" "64 9216 9216 5 2"; do
  echo "== $cfg split=2 pk=2 poly=2"
  GGML_B200_ATTN_PK=2 GGML_B200_ATTN_POLY=2 timeout 60 $A $cfg 0 2>&1 | grep -v "^softmax\|^blk\|^mma\|^clock"
  for v in "2 0" "2 1" "2 2" "2 3" "1 0" "1 1" "1 2" "1 3"; do
    set -- $v
    echo "== $cfg split=5 pk=$1 poly=$2"
    GGML_B200_ATTN_SPLIT=5 GGML_B200_ATTN_PK=$1 GGML_B200_ATTN_POLY=$2 timeout 60 $A $cfg 0 2>&1 | grep -v "^softmax\|^blk\|^mma\|^clock"
  done
done
for cfg in "40 1024 1024 8 16" "40 4096 4000 8 2" "64 1000 1090 3 2" "48 300 200 2 1" "40 300 4096 8 2" "64 128 256 1 1" "64 129 257 1 1" "64 1024 191 2 1" "40 1024 193 2 1" "32 512 512 4 2" "16 640 320 2 2" "40 256 129 2 2" "40 256 320 2 2" "64 256 321 2 2"; do
  for v in "2 2" "1 3"; do
    set -- $v
    echo "== $cfg split=5 pk=$1 poly=$2"
    GGML_B200_ATTN_SPLIT=5 GGML_B200_ATTN_PK=$1 GGML_B200_ATTN_POLY=$2 timeout 60 $A $cfg 0 2>&1 | grep -v "^softmax\|^blk\|^mma\|^clock"
  done
done
} > gpurun_out/attn_$TAG.log 2>&1
grep "^==\|us \|max abs" gpurun_out/attn_$TAG.log | paste - - - | awk '{print $2,$3,$4,$5,$6,$7,$8,$9, $16, $17, $NF}'
