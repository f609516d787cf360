#!/bin/bash
# one-wave CTA-pair tiling instead of split-K for the 8x8-level convolutions: tests, the two shapes, the UNet evaluation
TAG=${1:-r3k}
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_parity_r2_gpu.py -m gpu -q -x 2>&1 | tail -2
timeout 120 python tools/gemm_bench.py conv:8,8,1280,1280,16 conv:8,8,2560,1280,16 cgn:8,8,1280,1280,16 2>&1 | grep -v "^\[ggml"
GGML_B200_GEMM_PAIRWAVE=0 timeout 120 python tools/gemm_bench.py conv:8,8,1280,1280,16 conv:8,8,2560,1280,16 2>&1 | grep -v "^\[ggml"
for e in 1 0; do
  echo "== PAIRWAVE=$e"
  GGML_B200_GEMM_PAIRWAVE=$e timeout 600 python tools/time_unet.py 16 sd1 2>&1 | tail -1
done
} > gpurun_out/pairwave_$TAG.log 2>&1
cat gpurun_out/pairwave_$TAG.log
