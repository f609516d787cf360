#!/bin/bash
# single-GPU check after the GroupNorm epilogue fusion: GPU suite, bench line, SDXL level-2 attention shape on both kernels
TAG=${1:-r3d}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s -rs -x > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_$TAG.log
grep -a "passed\|failed\|^FAILED" gpurun_out/pytest_$TAG.log | tail -5
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$TAG.json"))
print("value %.2f e2e %.2f ms/step %.1f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), d["clocks"])
print({k:v for k,v in d["roofline"].items() if k in ("achieved","frac") or "unet" in k or "attention" in k})
print("sdxl", {k:v for k,v in d["config"]["sdxl_1024"].items() if k in ("images_per_sec","e2e_images_per_sec","unet_eval_ms_batch4","unet_frac_of_peak")})
print("vae", d["vae"])
print("hbm", {k:(round(v["frac"],3), round(v["ms"],3)) for k,v in d["roofline_hbm"]["families"].items()}, d["roofline_hbm"]["share_of_unet_eval"])
PY
A=mlimgsynth_b200/build/attn_trace
for cfg in "64 1024 1024 20 4" "64 4096 4096 10 4" "64 2304 2304 10 2" "64 9216 9216 5 2"; do
  for sp in 2 5; do
    echo "== $cfg split=$sp"; GGML_B200_ATTN_SPLIT=$sp timeout 60 $A $cfg 0 2>&1 | grep "us \|max abs"
  done
done > gpurun_out/attn_$TAG.log 2>&1
paste - - - < gpurun_out/attn_$TAG.log | awk '{print $2,$3,$4,$5,$6,$7, $14, $15, $NF}'
