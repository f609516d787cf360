#!/bin/bash
# attn_ap_kernel with the deferred hand-over: timing on the benchmark shapes + correctness on edge shapes
TAG=${1:-r2y}
mkdir -p gpurun_out
A=mlimgsynth_b200/build/attn_trace
run() { GGML_B200_ATTN_SPLIT=5 GGML_B200_ATTN_PK=$1 GGML_B200_ATTN_POLY=$2 timeout 30 $A $3 0 2>&1 | grep -v "^softmax\|^blk\|^mma\|^clock" | grep -v "thread [1-9]" ; }
{
echo "== 40 1024 1024 8 16 split=5 pk=2 poly=2"; run 2 2 "40 1024 1024 8 16"
if grep -q "timed out\|error" gpurun_out/attn_$TAG.log; then echo "ABORT: first case failed"; exit 1; fi
for cfg in "40 4096 4096 8 16" "64 4096 4096 10 4" "64 9216 9216 5 2"; do
  for v in "2 0" "2 1" "2 2" "1 1"; do
    set -- $v
    echo "== $cfg split=5 pk=$1 poly=$2"; run $1 $2 "$cfg"
  done
done
for cfg in "40 4096 4000 8 2" "64 1000 1090 3 2" "48 300 200 2 1" "40 300 4096 8 2" "64 128 256 1 1" "64 129 257 1 1" "64 1024 191 2 1" "40 1024 193 2 1" "32 512 512 4 2" "16 640 320 2 2" "40 256 129 2 2" "40 256 320 2 2" "64 256 321 2 2" "40 256 384 2 2" "40 256 448 2 2"; do
  for v in "2 1" "1 3"; do
    set -- $v
    echo "== $cfg split=5 pk=$1 poly=$2"; run $1 $2 "$cfg"
  done
done
} > gpurun_out/attn_$TAG.log 2>&1
grep "^==\|us \|max abs\|timed out\|ABORT" gpurun_out/attn_$TAG.log | awk '/^==/{if (h) print h, r; h=$0; r=""} !/^==/{r=r" | "$0} END{print h, r}' | sed 's/d=[0-9]* nq=[0-9]* nk=[0-9]* H=[0-9]* B=[0-9]* ://; s/[0-9.]* TFLOP.s//; s/max abs err vs f64 reference on 3 rows://' | cut -c1-150
