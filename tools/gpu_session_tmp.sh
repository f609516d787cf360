mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench15.json 2> gpurun_out/bench15.err; python -c "
import json; d=json.load(open('gpurun_out/bench15.json')); print(d['value'], d['e2e']['value'], d['roofline']['unet_eval_ms_batch16'], d['roofline']['frac'], {k:round(v['ms'],3) for k,v in d['kernel_profile'].items()}, d['config']['sdxl_1024'])"
