#!/bin/bash
# Round-2 first run: whole GPU suite (incl. the new full-shape parity, tile split, Level-2 drop-in tests), bench line, step profile.
TAG=${1:-r2a}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -rs --durations=15 > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_$TAG.log
tail -40 gpurun_out/pytest_$TAG.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"
tail -c 1500 gpurun_out/bench_$TAG.err
GGML_B200_PROFILE_STEPS=1 timeout 300 python tools/profile_unet.py 16 > gpurun_out/steps_$TAG.log 2>&1
ls -la gpurun_out | grep $TAG
