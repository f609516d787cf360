#!/bin/bash
TAG=${1:-r2c}
mkdir -p gpurun_out
timeout 60 mlimgsynth_b200/build/tmem_rates > gpurun_out/tmem_rates_$TAG.log 2>&1; cat gpurun_out/tmem_rates_$TAG.log
A=mlimgsynth_b200/build/attn_trace
GGML_B200_ATTN_SPLIT=1 timeout 120 $A 40 4096 4096 8 16 10 > gpurun_out/attn_trace_split_$TAG.log 2>&1; cat gpurun_out/attn_trace_split_$TAG.log
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_l2_dropin_gpu.py tests/test_vae_tiles_gpu.py -m gpu -q -x 2>&1 | tail -5
GGML_B200_PROFILE_STEPS=1 timeout 300 python tools/profile_unet.py 16 2> gpurun_out/steps_$TAG.log; python tools/summarize_steps.py gpurun_out/steps_$TAG.log | head -12
