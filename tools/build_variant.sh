#!/bin/bash
# Build an A/B variant of the library: tools/build_variant.sh <name> "<extra nvcc flags for warp_fuse.cu>"
# -> jafpro_b200/libjafpro_b200_<name>.so ; select it with JAFPRO_B200_LIB=... (see jafpro_b200/_lib.py)
set -e
cd "$(dirname "$0")/../jafpro_b200/csrc"
make -j8 > /dev/null
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -I../../include -I. --expt-relaxed-constexpr $2 -c warp_fuse.cu -o _obj/warp_fuse_$1.o
objs=$(ls _obj/*.o | grep -v "warp_fuse")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libjafpro_b200_$1.so $objs _obj/warp_fuse_$1.o
echo built ../libjafpro_b200_$1.so
