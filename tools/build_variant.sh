#!/bin/bash
# tuning experiment: tools/build_variant.sh NAME [-DMACRO=VALUE ...] -> tools/_variant_NAME.so (not committed)
set -e
cd "$(dirname "$0")/../rain_rendering_b200/csrc"
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC,-ffp-contract=off \
     -shared -cudart static "$@" -o ../../tools/_variant_$name.so rr_api.cu rr_kernels.cu rr_png_gpu.cu rr_sim.cu rr_host.cpp rr_host_xml.cpp rr_host_png.cpp -lz
echo tools/_variant_$name.so
