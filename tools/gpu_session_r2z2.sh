#!/bin/bash
# after making the GroupNorm-statistics epilogue a compile-time variant: op tests, bench line, launch list + GEMM traffic re-captured
TAG=${1:-r2z2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_parity_r2_gpu.py -m gpu -q -x 2>&1 | tail -2
timeout 300 python tools/time_unet.py 16 sd1 2>&1 | tail -1
timeout 300 python tools/time_unet.py 4 sdxl 2>&1 | tail -1
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"
L=$(python -c "import json;d=json.load(open('gpurun_out/bench_$TAG.json'));print(d['gpu_launches']//d['steps'])")
G=$(python -c "import json;d=json.load(open('gpurun_out/bench_$TAG.json'));print(d['roofline']['launches_per_unet_eval'])")
echo "launches per generation: $L, gemm launches per UNet evaluation: $G"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$TAG.json"))
print("value %.2f e2e %.2f ms/step %.1f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), d["clocks"])
print({k:v for k,v in d["roofline"].items() if k in ("achieved","frac") or "unet" in k or "attention" in k})
print("sdxl", {k:v for k,v in d["config"]["sdxl_1024"].items() if k in ("images_per_sec","e2e_images_per_sec","unet_eval_ms_batch4","unet_frac_of_peak")})
PY
GGML_B200_PROFILE_STEPS=1 timeout 300 python tools/profile_unet.py 16 > gpurun_out/steps_$TAG.log 2>&1
GGML_B200_PROFILE_STEPS=1 timeout 300 python tools/profile_unet.py 4 sdxl > gpurun_out/steps_sdxl_$TAG.log 2>&1
GGML_B200_NO_CUDA_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip $((2*L+2000)) -c $L --csv \
  --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sdxl > gpurun_out/bench_under_ncu_$TAG.log 2>&1
GGML_B200_NO_CUDA_GRAPH=1 timeout 600 ncu --clock-control none -k regex:gemm_tc --launch-skip $((2*G)) -c $G --csv \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,launch__grid_size \
  --log-file gpurun_out/gemm_traffic_$TAG.csv python tools/profile_unet.py 16 > gpurun_out/ncu_gemm_traffic_$TAG.log 2>&1
