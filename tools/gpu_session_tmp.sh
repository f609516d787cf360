mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -x -q > gpurun_out/pytest_ops4.log 2>&1; tail -3 gpurun_out/pytest_ops4.log
python tools/gemm_bench.py geglu:65536,320,1280 geglu:16384,640,2560 geglu:4096,1280,5120 res:65536,320,320 65536,320,320 65536,960,320 16384,640,640 4096,1280,1280 65536,320,1280 conv:64,64,320,320,16 conv:32,32,640,640,16 conv:16,16,1280,1280,16 conv:8,8,1280,1280,16 2>&1 | grep -v "^\[ggml" > gpurun_out/opbench4.log; cat gpurun_out/opbench4.log
GGML_B200_GEMM_TRACE=1 python tools/gemm_bench.py geglu:65536,320,1280 > gpurun_out/gemm_trace_geglu4.log 2>&1
GGML_B200_GEMM_TRACE=1 python tools/gemm_bench.py 65536,320,320 > gpurun_out/gemm_trace_lin4.log 2>&1
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench4.json 2> gpurun_out/bench4.err; python -c "
import json; d=json.load(open('gpurun_out/bench4.json')); print(d['value'], d['roofline']['unet_eval_ms_batch16'], {k:round(v['ms'],3) for k,v in d['kernel_profile'].items()}, d['config']['sdxl_1024'])"
