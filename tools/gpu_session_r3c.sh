#!/bin/bash
# GroupNorm statistics in the GEMM epilogue: op tests, then the UNet evaluation time with and without
TAG=${1:-r3c}
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -q -x 2>&1 | tail -5
for e in 0 1; do
  echo "== NO_GN_EPILOGUE=$e"
  GGML_B200_NO_GN_EPILOGUE=$e timeout 600 python tools/time_unet.py 16 sd1 2>&1 | tail -1
  GGML_B200_NO_GN_EPILOGUE=$e timeout 600 python tools/time_unet.py 4 sdxl 2>&1 | tail -1
done
} > gpurun_out/gnepi_$TAG.log 2>&1
cat gpurun_out/gnepi_$TAG.log
