#!/bin/bash
# Round-2 closing evidence run (one GPU): whole GPU suite, smoke, the default bench line, per-step profiles, ncu launch list of the
# bench command, DRAM traffic + tensor-pipe activity of every GEMM/conv launch of one UNet evaluation, full ncu captures of the
# dominant kernels (two-tile attention kernel, key-halves kernel, conv + GroupNorm-statistics epilogue, apply-only GroupNorm).
TAG=${1:-r2z}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/smi_$TAG.txt
timeout 1500 python -m pytest tests -m gpu -q -s -rs > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_$TAG.log
grep -a "passed\|failed\|^FAILED" gpurun_out/pytest_$TAG.log | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"
L=$(python -c "import json;d=json.load(open('gpurun_out/bench_$TAG.json'));print(d['gpu_launches']//d['steps'])")
G=$(python -c "import json;d=json.load(open('gpurun_out/bench_$TAG.json'));print(d['roofline']['launches_per_unet_eval'])")
echo "launches per generation: $L, gemm launches per UNet evaluation: $G"
GGML_B200_PROFILE_STEPS=1 timeout 300 python tools/profile_unet.py 16 > gpurun_out/steps_$TAG.log 2>&1
GGML_B200_PROFILE_STEPS=1 timeout 300 python tools/profile_unet.py 4 sdxl > gpurun_out/steps_sdxl_$TAG.log 2>&1
GGML_B200_PROFILE_STEPS=1 timeout 300 python tools/profile_vae.py > gpurun_out/steps_vae_$TAG.log 2>&1
timeout 300 python tools/time_unet.py 16 sd1 2>&1 | tail -1
timeout 300 python tools/time_unet.py 4 sdxl 2>&1 | tail -1
GGML_B200_NO_CUDA_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip $((2*L+2000)) -c $L --csv \
  --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sdxl > gpurun_out/bench_under_ncu_$TAG.log 2>&1
GGML_B200_NO_CUDA_GRAPH=1 timeout 600 ncu --clock-control none -k regex:gemm_tc --launch-skip $((2*G)) -c $G --csv \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,launch__grid_size \
  --log-file gpurun_out/gemm_traffic_$TAG.csv python tools/profile_unet.py 16 > gpurun_out/ncu_gemm_traffic_$TAG.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_ap_kernel --launch-skip 3 -c 1 -o gpurun_out/attn_ap_$TAG -f mlimgsynth_b200/build/attn_trace 40 4096 4096 8 16 0 > gpurun_out/ncu_attn_ap_$TAG.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_split_kernel --launch-skip 3 -c 1 -o gpurun_out/attn_self64_$TAG -f mlimgsynth_b200/build/attn_trace 64 4096 4096 10 4 0 > gpurun_out/ncu_attn_self64_$TAG.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:(gemm_tc_persistent|gn_apply_fast)' --launch-skip 9 -c 3 -o gpurun_out/gn320_$TAG -f python tools/gemm_bench.py gn:64,64,320,16 > gpurun_out/ncu_gn320_$TAG.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_persistent --launch-skip 3 -c 1 -o gpurun_out/conv320_$TAG -f python tools/gemm_bench.py conv:64,64,320,320,16 > gpurun_out/ncu_conv320_$TAG.log 2>&1
ls -la gpurun_out | grep $TAG
