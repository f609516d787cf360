#!/bin/bash
# Round-1 evidence run (session d): parity tests, bench line, ncu launch list of the bench command, DRAM traffic and
# tensor-pipe activity of EVERY GEMM/conv launch of one UNet evaluation, full ncu captures of the dominant shapes.
TAG=${1:-r1d}
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/smi_$TAG.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_$TAG.log
tail -3 gpurun_out/pytest_$TAG.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"
L=$(python -c "import json;d=json.load(open('gpurun_out/bench_$TAG.json'));print(d['gpu_launches']//d['steps'])")
G=$(python -c "import json;d=json.load(open('gpurun_out/bench_$TAG.json'));print(d['roofline']['launches_per_unet_eval'])")
echo "launches per generation: $L, gemm launches per UNet evaluation: $G"
# per-step CUDA-event times of one UNet evaluation (eager): the shape-by-shape picture
GGML_B200_PROFILE_STEPS=1 timeout 300 python tools/profile_unet.py 16 > gpurun_out/steps_$TAG.log 2>&1
GGML_B200_PROFILE_STEPS=1 timeout 300 python tools/profile_unet.py 4 sdxl > gpurun_out/steps_sdxl_$TAG.log 2>&1
# launch list: one generation's worth of consecutive launches (cyclic window) of the bench command, eager mode
GGML_B200_NO_CUDA_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip $((2*L+2000)) -c $L --csv \
  --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sdxl > gpurun_out/bench_under_ncu_$TAG.log 2>&1
# DRAM bytes + tensor pipe of every GEMM/conv launch of one UNet evaluation (the roofline's `traffic`)
GGML_B200_NO_CUDA_GRAPH=1 timeout 600 ncu --clock-control none -k regex:gemm_tc --launch-skip $((2*G)) -c $G --csv \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,launch__grid_size \
  --log-file gpurun_out/gemm_traffic_$TAG.csv python tools/profile_unet.py 16 > gpurun_out/ncu_gemm_traffic_$TAG.log 2>&1
# full captures: dominant conv / linear shapes of the SD1.5 batch-16 evaluation, the level-0 self-attention, the norms
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_persistent --launch-skip 3 -c 1 \
  -o gpurun_out/conv320_$TAG -f python tools/gemm_bench.py conv:64,64,320,320,16 > gpurun_out/ncu_conv320_$TAG.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_persistent --launch-skip 3 -c 1 \
  -o gpurun_out/conv640_$TAG -f python tools/gemm_bench.py conv:32,32,640,640,16 > gpurun_out/ncu_conv640_$TAG.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_persistent --launch-skip 3 -c 1 \
  -o gpurun_out/lin320_$TAG -f python tools/gemm_bench.py 65536,320,320 > gpurun_out/ncu_lin320_$TAG.log 2>&1
GGML_B200_NO_CUDA_GRAPH=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_tc_kernel -c 2 \
  -o gpurun_out/attn_$TAG -f python tools/profile_unet.py 16 > gpurun_out/ncu_attn_$TAG.log 2>&1
GGML_B200_NO_CUDA_GRAPH=1 timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:gn_(stats|apply)_fast|layernorm_fast' -c 6 \
  -o gpurun_out/norm_$TAG -f python tools/profile_unet.py 16 > gpurun_out/ncu_norm_$TAG.log 2>&1
ls -la gpurun_out | grep $TAG
