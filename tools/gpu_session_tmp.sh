mkdir -p gpurun_out
B=mlimgsynth_b200/build
for sh in "40 4096 77 8 16" "80 1024 77 8 16" "64 1024 77 20 4" "64 4096 77 10 4" "40 300 77 8 2" "128 1000 128 4 3" "64 130 16 2 1"; do
  for kv in 1 0; do echo -n "KV1=$kv: "; GGML_B200_ATTN_KV1=$kv timeout 60 $B/attn_trace $sh 1 2>&1 | grep -E "TFLOP|max abs" | tr '\n' ' '; echo; done
done
timeout 600 python -m pytest tests/test_ops_gpu.py -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench12.json 2> gpurun_out/bench12.err; python -c "
import json; d=json.load(open('gpurun_out/bench12.json')); print(d['value'], d['e2e']['value'], d['roofline']['unet_eval_ms_batch16'], d['roofline']['frac'], {k:round(v['ms'],3) for k,v in d['kernel_profile'].items()}, d['config']['sdxl_1024'])"
