#!/bin/bash
# Evidence run: bench line, ncu launch list of the bench command (one generation's worth of consecutive launches), and
# full ncu captures of the dominant kernels. Usage: tools/gpu_session_profiles.sh TAG
TAG=${1:-r1}
set -x
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
L=$(python -c "import json;d=json.load(open('gpurun_out/bench_$TAG.json'));print(d['gpu_launches']//d['steps'])")
echo "launches per generation: $L"
GGML_B200_NO_CUDA_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip $((2*L+2000)) -c $L --csv \
  --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_$TAG.log 2>&1
GGML_B200_NO_CUDA_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc_kernel -c 2 \
  -o gpurun_out/attn_$TAG -f python tools/profile_unet.py 16 > gpurun_out/ncu_attn_$TAG.log 2>&1
GGML_B200_NO_CUDA_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_persistent --launch-skip 4 -c 16 \
  -o gpurun_out/gemm_$TAG -f python tools/profile_unet.py 16 > gpurun_out/ncu_gemm_$TAG.log 2>&1
GGML_B200_NO_CUDA_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:gn_(stats|apply)_fast|layernorm_fast' -c 6 \
  -o gpurun_out/norm_$TAG -f python tools/profile_unet.py 16 > gpurun_out/ncu_norm_$TAG.log 2>&1
ls -la gpurun_out | grep $TAG
