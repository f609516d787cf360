#!/bin/bash
# 2-GPU session: NCCL tile gather test, c5 / c2 at N = 1, 2 (strong scaling), c1 e2e at N = 2 (device-side RGB8 gather).
TAG=${1:-r2m2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_vae_tiles_gpu.py tests/test_cfg_split_gpu.py -m gpu -q -s -k "nccl or cfg" 2>&1 | tail -8
for n in 1 2; do
  if [ $n = 1 ]; then L="python"; else L="$TR --nproc-per-node $n --master-port 2950$n"; fi
  timeout 600 $L bench.py --gpus $n --workload c5 --steps 5 --warmup 3 > gpurun_out/bench_c5_n${n}_$TAG.json 2> gpurun_out/bench_c5_n${n}_$TAG.err; echo "c5 n=$n exit $?"; tail -c 400 gpurun_out/bench_c5_n${n}_$TAG.err
  timeout 900 $L bench.py --gpus $n --workload c2 --steps 2 --warmup 3 > gpurun_out/bench_c2_n${n}_$TAG.json 2> gpurun_out/bench_c2_n${n}_$TAG.err; echo "c2 n=$n exit $?"; tail -c 400 gpurun_out/bench_c2_n${n}_$TAG.err
done
timeout 900 $TR --nproc-per-node 2 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --no-sdxl > gpurun_out/bench_c1_n2_$TAG.json 2> gpurun_out/bench_c1_n2_$TAG.err; echo "c1 n=2 exit $?"
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_c*_$TAG.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, "value %.3f e2e %.3f ms/step %.1f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), d["config"].get("tae"))
    except Exception as e: print(f, "ERR", e)
PY
