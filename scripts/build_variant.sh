#!/bin/bash
# tuning build: scripts/build_variant.sh <name> <file.cu to substitute for csrc/<basename>> [extra nvcc flags...]
# -> painter_b200/libpk_<name>.so (git-ignored; select at run time with PK_LIB=...)
set -e
name=$1; src=$2; shift 2
base=$(basename $src .cu)
o=/tmp/pk_variant_$name.o
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr \
  -Ipainter_b200/csrc "$@" -c $src -o $o
objs=$(ls painter_b200/build/*.o | grep -v "/$base.o")
nvcc -shared -o painter_b200/libpk_$name.so $objs $o -gencode arch=compute_100a,code=sm_100a
echo painter_b200/libpk_$name.so
