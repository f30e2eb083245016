#!/bin/bash
# tools/build_variant.sh <name> "<-D flags>" [source.cu]: one kernel source (default frontier_pack.cu) compiled with extra flags,
# linked with the current objects into bobrapet_b200/lib_ab/<name>.so (same-box A/B timing through BF_LIB)
set -e
cd "$(dirname "$0")/../bobrapet_b200/csrc"
SRC=${3:-frontier_pack.cu}
NV="/usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --cudart static -Xptxas -v"
mkdir -p ../lib_ab/obj
$NV $2 -c $SRC -o ../lib_ab/obj/$1.o 2> ../lib_ab/obj/$1.log
grep -h "Used\|spill" ../lib_ab/obj/$1.log | paste - - | sed 's/ptxas info    ://g' | head -4
objs=$(ls ../lib/obj/*.o | grep -v "$SRC.o")
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a --cudart static -shared ../lib_ab/obj/$1.o $objs -o ../lib_ab/$1.so
echo built lib_ab/$1.so
