#!/bin/bash
# GroupNorm kernels without FP64 in the block sums, cheaper apply set-up: tests + timing
TAG=${1:-r3i}
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "groupnorm or resnet or vae or transf" 2>&1 | tail -2
for e in 0 1; do
  echo "== NO_GN_EPILOGUE=$e"
  GGML_B200_NO_GN_EPILOGUE=$e timeout 300 python tools/gemm_bench.py cgn:64,64,320,320,16 cgn:32,32,640,640,16 cgn:16,16,1280,1280,16 2>&1 | grep -v "^\[ggml"
done
timeout 600 python tools/time_unet.py 16 sd1 2>&1 | tail -1
timeout 600 python tools/time_unet.py 4 sdxl 2>&1 | tail -1
} > gpurun_out/gn_$TAG.log 2>&1
cat gpurun_out/gn_$TAG.log
