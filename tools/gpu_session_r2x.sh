#!/bin/bash
# attn_ap_kernel: which resource bounds it? compile-time variants (tools/build_attn_variant.sh): dbg7 = no products, dbg8 = no exponentials
TAG=${1:-r2x}
mkdir -p gpurun_out
{
for cfg in "40 4096 4096 8 16" "64 4096 4096 10 4"; do
  for v in "" _dbg7 _dbg8; do
    for poly in 1 3; do
    echo "== $cfg variant=$v poly=$poly"
    GGML_B200_ATTN_SPLIT=5 GGML_B200_ATTN_PK=2 GGML_B200_ATTN_POLY=$poly timeout 30 mlimgsynth_b200/build/attn_trace$v $cfg 0 2>&1 | grep "us "
    done
  done
done
} > gpurun_out/attn_$TAG.log 2>&1
cat gpurun_out/attn_$TAG.log | paste - - | cut -c1-160
