#!/bin/bash
# Build a per-role timer variant of the grouped ConvLSTM kernels: tools/build_cgprof.sh [name] ["extra nvcc flags"]
# -> jafpro_b200/libjafpro_b200_<name>.so (default name: cgprof)
set -e
name=${1:-cgprof}
cd "$(dirname "$0")/../jafpro_b200/csrc"
make -j8 > /dev/null
mkdir -p _objv
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -I../../include -I. --expt-relaxed-constexpr -DJAF_GROUPED_PROFILE $2 -c convlstm_grouped.cu -o _objv/cg_$name.o
objs=$(ls _obj/*.o | grep -v convlstm_grouped)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libjafpro_b200_$name.so $objs _objv/cg_$name.o
echo built ../libjafpro_b200_$name.so
