#!/bin/bash
# single-GPU check: whole GPU suite, smoke, default bench line
TAG=${1:-r3a}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s -rs -x > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_$TAG.log
grep -a "passed\|failed\|^FAILED" gpurun_out/pytest_$TAG.log | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$TAG.json"))
print("value %.2f e2e %.2f ms/step %.1f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), d["clocks"])
print({k:v for k,v in d["roofline"].items() if k in ("achieved","frac") or "unet" in k or "attention" in k})
print("sdxl", {k:v for k,v in d["config"]["sdxl_1024"].items() if k in ("images_per_sec","e2e_images_per_sec","unet_eval_ms_batch4","unet_frac_of_peak")})
PY
