#!/bin/bash
# N-GPU session (N = $2, default 8): the weak-scaling headline (c1) at N GPUs after the closing kernel changes of round 2
TAG=${1:-r2m8b}; N=${2:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N"
python -c "import sys; sys.path.insert(0,'.'); import bench; [bench.weights_path(k) for k in ('sd1',)]"
timeout 900 $TR --master-port $((29600 + RANDOM % 300)) bench.py --gpus $N --steps 5 --warmup 3 --no-sdxl > gpurun_out/bench_c1_n${N}_$TAG.json 2> gpurun_out/bench_c1_n${N}_$TAG.err; echo "c1 n=$N exit $?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_c1_n${N}_$TAG.json").read().strip().splitlines()[-1])
print("value %.3f e2e %.3f ms/step %.1f scaling %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["scaling"]), d["clocks"])
PY
