#!/bin/bash
# tile / cluster / cta_group::2 sweep of the persistent GEMM on given shapes (kernel-only times)
# usage: gemm_sweep.sh "shape shape ..." cfg cfg ...     cfg = bn,cm,cn[,two_sm]
shapes=$1; shift
for shape in $shapes; do
  echo "== $shape"
  GGML_B200_GEMM_DEBUG=1 python tools/gemm_bench.py $shape 2>&1 | grep -E "kernel|BN=" | sed 's/|.*//' | tail -2
  for f in "$@"; do
    echo -n "force $f: "; GGML_B200_GEMM_FORCE=$f python tools/gemm_bench.py $shape 2>&1 | grep kernel | sed 's/|.*//'
  done
done
