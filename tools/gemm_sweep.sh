#!/bin/bash
# tile / cluster sweep of the persistent GEMM on the UNet's dominant shapes (kernel-only times)
for shape in conv:64,64,320,320,16 conv:16,16,1280,1280,16 65536,320,320 conv:8,8,1280,1280,16 16384,640,640; do
  echo "== $shape"
  GGML_B200_GEMM_DEBUG=1 python tools/gemm_bench.py $shape 2>&1 | grep -E "kernel|BN=" | tail -2
  for f in "$@"; do
    echo -n "force $f: "; GGML_B200_GEMM_FORCE=$f python tools/gemm_bench.py $shape 2>&1 | grep kernel | sed 's/|.*//'
  done
done
