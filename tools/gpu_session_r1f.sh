#!/bin/bash
# Round-1 closing run after the attention changes: parity tests, bench line, launch list, attention captures.
TAG=${1:-r1f}
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_$TAG.log
tail -3 gpurun_out/pytest_$TAG.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"
L=$(python -c "import json;d=json.load(open('gpurun_out/bench_$TAG.json'));print(d['gpu_launches']//d['steps'])")
GGML_B200_PROFILE_STEPS=1 timeout 300 python tools/profile_unet.py 16 > gpurun_out/steps_$TAG.log 2>&1
GGML_B200_PROFILE_STEPS=1 timeout 300 python tools/profile_unet.py 4 sdxl > gpurun_out/steps_sdxl_$TAG.log 2>&1
GGML_B200_NO_CUDA_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip $((2*L+2000)) -c $L --csv \
  --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sdxl > gpurun_out/bench_under_ncu_$TAG.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_tc_kernel --launch-skip 3 -c 1 -o gpurun_out/attn_self_$TAG -f mlimgsynth_b200/build/attn_trace 40 4096 4096 8 16 1 > gpurun_out/ncu_attn_self_$TAG.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_tc_kernel --launch-skip 3 -c 1 -o gpurun_out/attn_self64_$TAG -f mlimgsynth_b200/build/attn_trace 64 4096 4096 10 4 1 > gpurun_out/ncu_attn_self64_$TAG.log 2>&1
ls -la gpurun_out | grep $TAG
