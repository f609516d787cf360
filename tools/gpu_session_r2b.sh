#!/bin/bash
# Round-2 run b: key-split attention kernel (timing + correctness vs the dual form), MUFU micro-benchmark, whole GPU suite, bench.
TAG=${1:-r2b}
mkdir -p gpurun_out
A=mlimgsynth_b200/build/attn_trace
{
for cfg in "40 4096 4096 8 16" "64 4096 4096 10 4" "64 9216 9216 5 2" "40 4096 4000 8 2" "64 1000 1090 3 2" "48 300 200 2 1"; do
  for split in 1 0; do
    for poly in 1 0 2; do
      [ $split = 0 ] && [ $poly != 1 ] && continue
      echo "== $cfg split=$split poly=$poly"
      GGML_B200_ATTN_SPLIT=$split GGML_B200_ATTN_POLY=$poly timeout 120 $A $cfg 0 2>&1 | grep -v "^softmax\|^blk"
    done
  done
done
} > gpurun_out/attn_$TAG.log 2>&1
cat gpurun_out/attn_$TAG.log
timeout 120 mlimgsynth_b200/build/pipe_rates > gpurun_out/pipe_rates_$TAG.log 2>&1; cat gpurun_out/pipe_rates_$TAG.log
timeout 2400 python -m pytest tests -m gpu -q -rs -s > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_$TAG.log
grep -a "PARITY\|passed\|failed\|^FAILED\|differ" gpurun_out/pytest_$TAG.log | tail -40
timeout 900 python bench.py --steps 3 --warmup 3 --no-sdxl --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$TAG.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "unet ms", {k:v for k,v in d["roofline"].items() if "unet" in k or k=="frac" or "attention" in k})
print("hbm", json.dumps(d.get("roofline_hbm"))[:800]); print("vae", d.get("vae"))
PY
