#!/bin/bash
# builds a private copy of the library with the FPS phase trace compiled in and prints the per-phase cycles
set -e
cd "$(dirname "$0")/../co-occ_b200/csrc"
mkdir -p /tmp/fps_trace
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr \
    -DCOOCC_FPS_TRACE -shared -o /tmp/fps_trace/libfps_trace.so gsf_index.cu -lcudart
cd ../..
python tools/fps_trace.py /tmp/fps_trace/libfps_trace.so 200 200 16
python tools/fps_trace.py /tmp/fps_trace/libfps_trace.so 100 100 8
