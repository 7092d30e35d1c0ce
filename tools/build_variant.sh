#!/bin/bash
# usage: tools/build_variant.sh NAME "EXTRA nvcc flags"  -> build/variants/NAME.so (walk.cu/emit.cu rebuilt with the flags)
# Development aid: A/B timing of kernel variants on the GPU box (tools/run_variants.sh).
set -e
cd "$(dirname "$0")/../sigtk_b200/csrc"
name=$1; extra=$2
tmp=$(mktemp -d)
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --fmad=false -Werror cross-execution-space-call -Xcompiler -fPIC -Xptxas -v"
for f in walk emit; do $NV $extra -dc $f.cu -o $tmp/$f.o 2> $tmp/$f.log; done
grep -A2 "walk_chunks_kernelILi[01]" $tmp/walk.log | grep -E "registers|spill" | tr "\n" " "; echo
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC $(ls *.o | grep -v -e "^walk.o$" -e "^emit.o$" | tr "\n" " ") $tmp/walk.o $tmp/emit.o -o ../../build/variants/$name.so
rm -rf $tmp
