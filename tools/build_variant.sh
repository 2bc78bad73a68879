#!/bin/bash
# Builds an experimental variant of libvdbm_b200.so with extra nvcc flags into tools/build/ (git-ignored; ships to the
# GPU box with gpurun). Usage: tools/build_variant.sh <name> [-DVDBM_KBATCH=8 ...]
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p tools/build/$name
for f in vdbm_kernels vdbm_abi vdbm_group; do
  nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
       -I include "$@" -c vdb_mapping_b200/csrc/$f.cu -o tools/build/$name/$f.o &
done
wait
nvcc -shared -o tools/build/$name/libvdbm_b200.so tools/build/$name/vdbm_kernels.o tools/build/$name/vdbm_abi.o tools/build/$name/vdbm_group.o -gencode arch=compute_100a,code=sm_100a
echo tools/build/$name/libvdbm_b200.so
