#!/bin/bash
TAG=${1:-r2i}
mkdir -p gpurun_out
A=mlimgsynth_b200/build/attn_trace
{
for cfg in "40 4096 4096 8 16" "64 4096 4096 10 4" "64 9216 9216 5 2" "40 4096 4000 8 2" "64 1000 1090 3 2" "48 300 200 2 1" "40 333 384 2 1" "64 128 256 1 1"; do
  for stag in 1 0; do
    for poly in 1 2 0; do
      [ $stag = 0 ] && [ $poly != 1 ] && continue
      echo "== $cfg stagger=$stag poly=$poly"
      GGML_B200_ATTN_STAGGER=$stag GGML_B200_ATTN_POLY=$poly timeout 120 $A $cfg 0 2>&1 | grep -v "^softmax\|^blk\|^mma"
    done
  done
done
} > gpurun_out/attn_$TAG.log 2>&1
grep -A1 "^==" gpurun_out/attn_$TAG.log | grep -v "^--" | paste - - | awk '{print $2,$3,$4,$5,$6,$7,$8, $14, $15}'
grep "max abs err" gpurun_out/attn_$TAG.log | sort | uniq -c | sort -k6 -g | tail -3
timeout 600 python tools/phase_times.py 8 2>&1 | grep -v "^\[" | tee gpurun_out/phase_times_$TAG.log
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_parity_r2_gpu.py -m gpu -q -s 2>&1 | grep -a "passed\|failed\|FAILED\|Error" | tail -5
