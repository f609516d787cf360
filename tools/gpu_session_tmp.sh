mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_host_gpu.py -x -q > gpurun_out/pytest11.log 2>&1; tail -3 gpurun_out/pytest11.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-sdxl > gpurun_out/bench11.json 2> gpurun_out/bench11.err; python -c "
import json; d=json.load(open('gpurun_out/bench11.json')); print(d['value'], d['e2e']['value'], d['roofline']['unet_eval_ms_batch16'], d['roofline']['frac'], {k:round(v['ms'],3) for k,v in d['kernel_profile'].items()})"
timeout 300 python tools/profile_vae.py 2>&1 | grep -E "vae decode|upscale" | tail -4
