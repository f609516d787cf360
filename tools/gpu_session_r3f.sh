#!/bin/bash
# GroupNorm statistics epilogue without shared-memory atomics: tests + timing
TAG=${1:-r3f}
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "groupnorm or gemm_modes or resnet or conv" 2>&1 | tail -2
timeout 300 python tools/gemm_bench.py gn:64,64,320,16 gn:32,32,640,16 2>&1 | grep -v "^\[ggml"
for e in 0 1; do
  echo "== NO_GN_EPILOGUE=$e"
  GGML_B200_NO_GN_EPILOGUE=$e timeout 600 python tools/time_unet.py 16 sd1 2>&1 | tail -1
  GGML_B200_NO_GN_EPILOGUE=$e timeout 600 python tools/time_unet.py 4 sdxl 2>&1 | tail -1
done
GGML_B200_PROFILE_STEPS=1 timeout 300 python tools/profile_unet.py 16 > gpurun_out/steps_$TAG.log 2>&1
python tools/summarize_steps.py gpurun_out/steps_$TAG.log 6
} > gpurun_out/gnepi_$TAG.log 2>&1
cat gpurun_out/gnepi_$TAG.log
