#!/bin/bash
# A/B builds of one kernel header: build_variant.sh NAME path/to/vo_gridgemm.cuh -> generative-physics-informed-pde_b200/build/libNAME.so
# (copy of csrc with the header replaced; vo.cu recompiled, the other objects reused from the last full build)
set -e
NAME=$1; HDR=$2
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
PKG=$ROOT/generative-physics-informed-pde_b200
W=/tmp/variants/$NAME
rm -rf $W; mkdir -p $W/pkg $W/include
cp -r $PKG/csrc $W/pkg/csrc; cp $ROOT/include/gpde_b200.h $W/include/
cp $HDR $W/pkg/csrc/$(basename ${3:-vo_gridgemm.cuh})
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xptxas -v -c -o $W/vo.o $W/pkg/csrc/vo.cu > $W/ptxas.log 2>&1
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $PKG/build/lib$NAME.so $PKG/build/rom.o $W/vo.o $PKG/build/prolong.o $PKG/build/fom_cg.o
grep -A2 "vo_gridgemm_kernelILi256ELb0Edd" $W/ptxas.log | grep spill
echo built lib$NAME.so
