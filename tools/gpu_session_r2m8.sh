#!/bin/bash
# N-GPU session (N = $2, default 8): strong-scaling lines of configs 2, 3, 5 and the weak-scaling headline at N GPUs.
TAG=${1:-r2m8}; N=${2:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N"
python -c "import sys; sys.path.insert(0,'.'); import bench; [bench.weights_path(k) for k in ('sd1','sd2','sdxl','tae')]"
run() { name=$1; shift; timeout 900 $TR --master-port $((29600 + RANDOM % 300)) bench.py --gpus $N "$@" > gpurun_out/bench_${name}_n${N}_$TAG.json 2> gpurun_out/bench_${name}_n${N}_$TAG.err; echo "$name n=$N exit $?"; tail -c 300 gpurun_out/bench_${name}_n${N}_$TAG.err | grep -v OMP_NUM | tail -3; }
run c5 --workload c5 --steps 5 --warmup 3
run c2 --workload c2 --steps 3 --warmup 3
run c3 --workload c3 --steps 2 --warmup 3
if [ "$3" != "noc1" ]; then run c1 --steps 3 --warmup 3; fi
timeout 600 python -m pytest tests/test_vae_tiles_gpu.py -m gpu -q -s -k nccl 2>&1 | tail -3
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_c*_n${N}_$TAG.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, "value %.3f e2e %.3f ms/step %.1f scaling %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["scaling"]), (d["config"].get("sdxl_1024") or {}).get("images_per_sec"))
    except Exception as e: print(f, "ERR", e)
PY
