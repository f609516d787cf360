#!/bin/bash
# Round-2 evidence run (one GPU): whole GPU suite, the default bench line, per-step profiles, ncu launch list of the bench command,
# DRAM traffic + tensor-pipe activity of every GEMM/conv launch of one UNet evaluation, full ncu captures of the dominant kernels.
TAG=${1:-r2g}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/smi_$TAG.txt
timeout 1500 python -m pytest tests -m gpu -q -s -rs > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_$TAG.log
grep -a "passed\|failed\|^FAILED" gpurun_out/pytest_$TAG.log | tail -5
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"
timeout 900 python bench.py --workload c3 --steps 1 --warmup 3 > gpurun_out/bench_c3_n1_$TAG.json 2> gpurun_out/bench_c3_n1_$TAG.err; echo "c3 n=1 exit $?"
L=$(python -c "import json;d=json.load(open('gpurun_out/bench_$TAG.json'));print(d['gpu_launches']//d['steps'])")
G=$(python -c "import json;d=json.load(open('gpurun_out/bench_$TAG.json'));print(d['roofline']['launches_per_unet_eval'])")
echo "launches per generation: $L, gemm launches per UNet evaluation: $G"
GGML_B200_PROFILE_STEPS=1 timeout 300 python tools/profile_unet.py 16 > gpurun_out/steps_$TAG.log 2>&1
GGML_B200_PROFILE_STEPS=1 timeout 300 python tools/profile_unet.py 4 sdxl > gpurun_out/steps_sdxl_$TAG.log 2>&1
GGML_B200_PROFILE_STEPS=1 timeout 300 python tools/profile_vae.py > gpurun_out/steps_vae_$TAG.log 2>&1
GGML_B200_NO_CUDA_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip $((2*L+2000)) -c $L --csv \
  --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sdxl > gpurun_out/bench_under_ncu_$TAG.log 2>&1
GGML_B200_NO_CUDA_GRAPH=1 timeout 600 ncu --clock-control none -k regex:gemm_tc --launch-skip $((2*G)) -c $G --csv \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,launch__grid_size \
  --log-file gpurun_out/gemm_traffic_$TAG.csv python tools/profile_unet.py 16 > gpurun_out/ncu_gemm_traffic_$TAG.log 2>&1
cap() { name=$1; shift; k=$1; shift; timeout 400 ncu --set full --clock-control none --import-source on -k "$k" --launch-skip 3 -c 1 -o gpurun_out/${name}_$TAG -f "$@" > gpurun_out/ncu_${name}_$TAG.log 2>&1; }
cap conv320 regex:gemm_tc_persistent python tools/gemm_bench.py conv:64,64,320,320,16
cap lin320 regex:gemm_tc_persistent python tools/gemm_bench.py res:65536,320,320
timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:gn_(stats|apply)_fast' --launch-skip 6 -c 2 -o gpurun_out/gn320_$TAG -f python tools/gemm_bench.py gn:64,64,320,16 > gpurun_out/ncu_gn320_$TAG.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_split_kernel --launch-skip 3 -c 1 -o gpurun_out/attn_self_$TAG -f mlimgsynth_b200/build/attn_trace 40 4096 4096 8 16 1 > gpurun_out/ncu_attn_self_$TAG.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_split_kernel --launch-skip 3 -c 1 -o gpurun_out/attn_self64_$TAG -f mlimgsynth_b200/build/attn_trace 64 4096 4096 10 4 1 > gpurun_out/ncu_attn_self64_$TAG.log 2>&1
ls -la gpurun_out | grep $TAG
