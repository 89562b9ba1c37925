#!/bin/bash
# usage: tools/build_variant.sh <name> <extra nvcc flags...>   -> writes variants/lib_<name>.so (git-ignored; travels with gpurun)
set -e
name=$1; shift
mkdir -p variants/obj_$name
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math -ccbin /usr/bin/g++"
/usr/local/cuda/bin/nvcc $F "$@" -Xptxas -v -c honeybadger_b200/csrc/hb_kernels.cu -o variants/obj_$name/k.o 2> variants/obj_$name/ptxas.log &
/usr/local/cuda/bin/nvcc $F "$@" -c honeybadger_b200/csrc/hb_api.cu -o variants/obj_$name/a.o &
wait
/usr/local/cuda/bin/nvcc -shared -o variants/lib_$name.so variants/obj_$name/k.o variants/obj_$name/a.o -cudart static -ccbin /usr/bin/g++ 2>/dev/null
echo "$name: $(grep -A2 'norm3_fast_kernelILb1ELb1ELb0' variants/obj_$name/ptxas.log | grep -oE '[0-9]+ bytes spill stores|Used [0-9]+ registers' | tr '\n' ' ')"
