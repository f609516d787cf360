#!/bin/bash
# LayerNorm tile kernel: parity tests + timing against the lane-owns-channels kernel (GGML_B200_LN_MODE=2)
TAG=${1:-r3b}
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "layernorm or transf or clip" 2>&1 | tail -3
for mode in 0 2; do
  echo "== LN_MODE=$mode"
  GGML_B200_LN_MODE=$mode timeout 300 python tools/gemm_bench.py ln:4096,1280 ln:16384,640 ln:65536,320 ln:16384,1280 ln:4096,640 ln:1024,1280 ln:1232,768 2>&1 | grep -v "^\[ggml"
done
} > gpurun_out/ln_$TAG.log 2>&1
cat gpurun_out/ln_$TAG.log
