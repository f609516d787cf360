#!/bin/bash
# development aid: a copy of the engine with attn_tc.cu compiled under extra macros (timeline probes, timing experiments)
#   tools/build_attn_variant.sh <name> <nvcc -D flags...>   ->  mlimgsynth_b200/build/<name>/libggml_b200.so + build/attn_trace_<name>
set -e
NAME=$1; shift
D=mlimgsynth_b200/build/$NAME
mkdir -p $D
nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -Xcompiler -fPIC -I include -I mlimgsynth_b200/csrc "$@" -x cu -c mlimgsynth_b200/csrc/attn_tc.cu -o $D/attn_tc.o
nvcc -shared -o $D/libggml_b200.so $D/attn_tc.o $(ls mlimgsynth_b200/build/*.o | grep -v attn_tc) -gencode arch=compute_100a,code=sm_100a
nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O2 -I include -I mlimgsynth_b200/csrc tools/attn_trace.cu -L $D -lggml_b200 -Xlinker -rpath="\$ORIGIN/$NAME" -o mlimgsynth_b200/build/attn_trace_$NAME
