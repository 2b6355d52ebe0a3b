#!/bin/bash
# build a variant of libdxmcb200.so with extra nvcc defines: tools/build_variant.sh <name> -DX=1 ...   -> dxmclib_b200/variants/<name>.so
set -e
cd "$(dirname "$0")/.."
NAME=$1; shift
H=dxmclib_b200
mkdir -p $H/variants $H/build
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Iinclude -I$H/include -I$H/host -I$H/csrc "$@" -c $H/csrc/transport.cu -o $H/build/transport_$NAME.o
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $H/variants/$NAME.so $H/build/transport_$NAME.o $H/build/xrl_lite.cpp.o $H/build/matdb.cpp.o $H/build/scene_capi.cpp.o -ldl -lpthread -Xlinker -Bsymbolic
echo built $H/variants/$NAME.so
