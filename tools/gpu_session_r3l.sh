#!/bin/bash
# short-K linears (+ residual) of the transformer blocks: the cost model's tile choice against forced alternatives
TAG=${1:-r3l}
mkdir -p gpurun_out
{
for spec in res:65536,320,320 res:16384,640,640 res:4096,1280,1280; do
  echo "== $spec default"; GGML_B200_GEMM_DEBUG=1 timeout 120 python tools/gemm_bench.py $spec 2>&1 | grep "BN=\|kernel" | grep -v "M=.*K=64 \|N=64 " | tail -2
  for f in 64,1,1,1 96,1,1,1 128,1,1,1 160,1,1,1 256,1,1,1 64,1,1,0 128,1,1,0 160,1,1,0 256,1,1,0 160,2,1,0 128,1,2,0; do
    echo "== $spec force=$f"; GGML_B200_GEMM_FORCE=$f timeout 120 python tools/gemm_bench.py $spec 2>&1 | grep "kernel" | tail -1
  done
done
} > gpurun_out/gemm_$TAG.log 2>&1
grep -A2 "^==" gpurun_out/gemm_$TAG.log | grep -v "^--" | awk '/^==/{h=$0; next} /kernel/{print h, "|", $3, $4, $5, $6} /BN=/{print h, "|", $0}' | cut -c1-170
