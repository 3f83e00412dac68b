#!/bin/bash
# usage: scripts/build_variant.sh <name> <extra nvcc flags...>   ->  binary-networks-pytorch_b200/csrc/variants/libbnn_b200_<name>.so
# A second build of the library with extra -D flags for kernel A/B runs (select it with BNN_B200_LIB=<path>).
set -e
NAME=$1; shift
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=$ROOT/binary-networks-pytorch_b200/csrc
TMP=$(mktemp -d)
mkdir -p $TMP/pkg/csrc $TMP/include $SRC/variants
cp $SRC/*.cu $SRC/*.cuh $SRC/Makefile $TMP/pkg/csrc/
cp $ROOT/include/*.h $TMP/include/
make -C $TMP/pkg/csrc -j8 CXXFLAGS="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xptxas -v $*" > $TMP/build.log 2>&1 || (tail -20 $TMP/build.log; exit 1)
cp $TMP/pkg/csrc/libbnn_b200.so $SRC/variants/libbnn_b200_$NAME.so
rm -rf $TMP
echo $SRC/variants/libbnn_b200_$NAME.so
