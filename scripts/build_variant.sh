#!/bin/bash
# usage: build_variant.sh NAME [git-ref] [extra nvcc flags...]: builds gams_b200/csrc/build/variants/lib_NAME.so with
# gc_wave_fused.cu taken from git-ref (default: working tree), for A/B runs on one GPU box (GAMS_CUDA_LIB=...).
# Every other object comes from the default build (make first).
set -e
cd "$(dirname "$0")/../gams_b200/csrc"
name=$1; ref=${2:-WORK}; shift; shift || true
mkdir -p build/variants /tmp/variant_$name
if [ "$ref" = WORK ]; then cp gc_wave_fused.cu /tmp/variant_$name/gc_wave_fused.cu; else git show $ref:gams_b200/csrc/gc_wave_fused.cu > /tmp/variant_$name/gc_wave_fused.cu; fi
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden -fmad=false -Xptxas -v -I. "$@" -c /tmp/variant_$name/gc_wave_fused.cu -o /tmp/variant_$name/fw.o 2> /tmp/variant_$name/ptxas.log
others=$(ls build/*.o | grep -v gc_wave_fused.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/lib_$name.so $others /tmp/variant_$name/fw.o -Xlinker --exclude-libs,ALL
echo built build/variants/lib_$name.so
