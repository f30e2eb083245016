#!/bin/bash
# tools/build_variant.sh <name> "<-D flags>": frontier_kernel.cu compiled with extra flags, linked with the current objects
# into bobrapet_b200/lib_ab/<name>.so (same-box A/B timing through BF_LIB)
set -e
cd "$(dirname "$0")/../bobrapet_b200/csrc"
NV="/usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --cudart static"
mkdir -p ../lib_ab/obj
$NV $2 -c frontier_kernel.cu -o ../lib_ab/obj/$1.o
objs=$(ls ../lib/obj/*.o | grep -v frontier_kernel.cu.o)
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a --cudart static -shared ../lib_ab/obj/$1.o $objs -o ../lib_ab/$1.so
echo built lib_ab/$1.so
