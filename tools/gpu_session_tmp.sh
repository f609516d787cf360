mkdir -p gpurun_out
python tools/sdxl_time.py 2>&1 | tail -6
timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -k "groupnorm" 2>&1 | tail -2
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-sdxl > gpurun_out/bench8.json 2> gpurun_out/bench8.err; python -c "
import json; d=json.load(open('gpurun_out/bench8.json')); print(d['value'], d['e2e']['value'], d['roofline']['unet_eval_ms_batch16'], d['roofline']['frac'], {k:round(v['ms'],3) for k,v in d['kernel_profile'].items()})"
