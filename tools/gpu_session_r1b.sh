#!/bin/bash
# Round-1 GPU session: parity tests, bench line, ncu launch list of the bench command, full captures of the top kernels.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r1b.log 2>&1
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1b.json 2> gpurun_out/bench_r1b.err
L=$(python -c "import json;d=json.load(open('gpurun_out/bench_r1b.json'));print(d['gpu_launches']//d['steps'])")
echo "launches per generation: $L"
# launch list: one full generation's worth of consecutive launches (cyclic window) of the bench command, eager mode
GGML_B200_NO_CUDA_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip $((2*L+3000)) -c $L --csv \
  --log-file gpurun_out/launches_r1b.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
# full captures
GGML_B200_NO_CUDA_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc_kernel -c 2 \
  -o gpurun_out/attn_r1b -f python tools/profile_unet.py 16 > gpurun_out/ncu_attn.log 2>&1
GGML_B200_NO_CUDA_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -c 14 \
  -o gpurun_out/gemm_r1b -f python tools/profile_unet.py 16 > gpurun_out/ncu_gemm.log 2>&1
GGML_B200_NO_CUDA_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:groupnorm -c 4 \
  -o gpurun_out/gn_r1b -f python tools/profile_unet.py 16 > gpurun_out/ncu_gn.log 2>&1
GGML_B200_PROFILE_STEPS=1 python tools/profile_unet.py 16 > gpurun_out/steps_r1b.log 2>&1
ls -la gpurun_out
