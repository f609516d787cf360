#!/bin/bash
TAG=${1:-r2f}
mkdir -p gpurun_out
A=mlimgsynth_b200/build/attn_trace
{
for cfg in "40 4096 4096 8 16" "64 4096 4096 10 4" "64 9216 9216 5 2" "40 4096 4000 8 2" "64 1000 1090 3 2" "48 300 200 2 1" "40 333 384 2 1" "64 128 256 1 1"; do
  for split in 4 2; do
    for poly in 1 2 0; do
      [ $split = 2 ] && [ $poly != 1 ] && continue
      echo "== $cfg split=$split poly=$poly"
      GGML_B200_ATTN_SPLIT=$split GGML_B200_ATTN_POLY=$poly timeout 120 $A $cfg 0 2>&1 | grep -v "^softmax\|^blk\|^mma"
    done
  done
done
} > gpurun_out/attn_$TAG.log 2>&1
cat gpurun_out/attn_$TAG.log
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_parity_r2_gpu.py tests/test_cfg_split_gpu.py tests/test_l2_dropin_gpu.py -m gpu -q -s 2>&1 | grep -a "PARITY\|passed\|failed\|FAILED\|CFG split\|Error" | tail -30
GGML_B200_PROFILE_STEPS=1 timeout 300 python tools/profile_unet.py 16 2> gpurun_out/steps_$TAG.log; python tools/summarize_steps.py gpurun_out/steps_$TAG.log 2>/dev/null | head -16
