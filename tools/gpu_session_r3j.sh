#!/bin/bash
# 8x8-level convolutions (M = 1024, N = 1280, K = 11520 / 23040): split-K kernel against CTA-pair tiles of one wave
TAG=${1:-r3j}
mkdir -p gpurun_out
{
for spec in conv:8,8,1280,1280,16 conv:8,8,2560,1280,16; do
  echo "== $spec default (split-K)"; timeout 120 python tools/gemm_bench.py $spec 2>&1 | grep -v "^\[ggml"
  for f in 80,1,1,1 96,1,1,1 128,1,1,1 64,1,1,1 160,1,1,1 256,1,1,1 128,1,1,0 64,2,1,0; do
    echo "== $spec splitk=0 force=$f"; GGML_B200_GEMM_SPLITK=0 GGML_B200_GEMM_FORCE=$f timeout 120 python tools/gemm_bench.py $spec 2>&1 | grep -v "^\[ggml"
  done
done
} > gpurun_out/gemm_$TAG.log 2>&1
paste - - < gpurun_out/gemm_$TAG.log | awk '{print $2,$3,$4,$5, $7,$8,$9,$10,$11}'
