#!/bin/bash
# what the GroupNorm-statistics epilogue costs the producing convolution: conv -> group_norm pairs with and without, epilogue timeline
TAG=${1:-r3g}
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "groupnorm or resnet" 2>&1 | tail -2
for e in 0 1; do
  echo "== NO_GN_EPILOGUE=$e"
  GGML_B200_NO_GN_EPILOGUE=$e timeout 300 python tools/gemm_bench.py cgn:64,64,320,320,16 cgn:32,32,640,640,16 cgn:16,16,1280,1280,16 cgn:64,64,640,320,16 2>&1 | grep -v "^\[ggml"
done
echo "== trace NO_GN_EPILOGUE=0"
GGML_B200_GEMM_TRACE=1 timeout 300 python tools/gemm_bench.py cgn:64,64,320,320,16 2>&1 | grep -A 9 "epilogue timeline" | head -10
for e in 0 1; do
  echo "== NO_GN_EPILOGUE=$e"
  GGML_B200_NO_GN_EPILOGUE=$e timeout 600 python tools/time_unet.py 16 sd1 2>&1 | tail -1
done
} > gpurun_out/gnepi_$TAG.log 2>&1
cat gpurun_out/gnepi_$TAG.log
