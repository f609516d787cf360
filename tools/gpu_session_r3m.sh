#!/bin/bash
# cross-attention kernel (one key block) with two CTAs per SM for the 77-token context and heads up to 48 wide
TAG=${1:-r3m}
mkdir -p gpurun_out
A=mlimgsynth_b200/build/attn_trace
{
for cfg in "40 4096 77 8 16" "40 1024 77 8 16" "40 300 77 8 2" "48 4096 80 8 2" "64 4096 77 10 4" "80 1024 77 8 16" "40 4096 100 8 2" "40 4096 17 8 2"; do
  echo "== $cfg"; timeout 60 $A $cfg 0 2>&1 | grep "us \|max abs\|error\|timed out" | head -3
done
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "attention or transf or clip" 2>&1 | tail -2
timeout 600 python tools/time_unet.py 16 sd1 2>&1 | tail -1
} > gpurun_out/kv1_$TAG.log 2>&1
cat gpurun_out/kv1_$TAG.log
