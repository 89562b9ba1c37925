#!/bin/bash
# usage: tools/build_variant.sh <name> <extra nvcc flags...>   -> gpurun_out is NOT used; writes variants/lib_<name>.so
set -e
name=$1; shift
mkdir -p variants/obj_$name
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math -ccbin /usr/bin/g++"
/usr/local/cuda/bin/nvcc $F "$@" -c honeybadger_b200/csrc/hb_kernels.cu -o variants/obj_$name/k.o &
/usr/local/cuda/bin/nvcc $F "$@" -c honeybadger_b200/csrc/hb_api.cu -o variants/obj_$name/a.o &
wait
/usr/local/cuda/bin/nvcc -shared -o variants/lib_$name.so variants/obj_$name/k.o variants/obj_$name/a.o -cudart static -ccbin /usr/bin/g++ 2>/dev/null
cuobjdump -res-usage variants/lib_$name.so 2>/dev/null | grep -A1 "norm3_kernelILb1ELb1ELb1ELb0" | grep -o "REG:[0-9]* STACK:[0-9]*"
