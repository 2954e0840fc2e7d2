#!/bin/sh
# developer build with the in-kernel clocks / task trace (-DB2_TIMING): scripts/_dev/libb2_timing.so
cd "$(dirname "$0")/../../cannoles_b200/csrc" && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
  -DB2_TIMING -Xcompiler -fPIC,-O3 -shared -o ../../scripts/_dev/libb2_timing.so engine.cu capi.cu batched.cu measure.cu \
  symbolic.cpp ordering.cpp /usr/local/cuda/lib64/libmetis_static.a -lcublas -Xlinker -rpath=/usr/local/cuda/lib64
