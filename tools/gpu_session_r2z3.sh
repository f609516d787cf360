#!/bin/bash
# launch list + GEMM traffic re-captured on the final code of round 2 (one wave of CTA pairs instead of split-K: 198 GEMM launches per evaluation)
TAG=${1:-r2z3}; L=${2:-7888}; G=${3:-198}
mkdir -p gpurun_out
GGML_B200_PROFILE_STEPS=1 timeout 300 python tools/profile_unet.py 16 > gpurun_out/steps_$TAG.log 2>&1
GGML_B200_NO_CUDA_GRAPH=1 timeout 600 ncu --clock-control none -k regex:gemm_tc --launch-skip $((2*G)) -c $G --csv \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,launch__grid_size \
  --log-file gpurun_out/gemm_traffic_$TAG.csv python tools/profile_unet.py 16 > gpurun_out/ncu_gemm_traffic_$TAG.log 2>&1
GGML_B200_NO_CUDA_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip $((2*L+2000)) -c $L --csv \
  --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sdxl > gpurun_out/bench_under_ncu_$TAG.log 2>&1
ls -la gpurun_out | grep $TAG
