#!/bin/bash
# usage: build_variant.sh NAME [git-ref] [extra nvcc flags...]: builds gams_b200/csrc/build/variants/lib_NAME.so with
# gc_wave_fused.cu taken from git-ref (default: working tree), for A/B runs on one GPU box (GAMS_CUDA_LIB=...).
set -e
cd "$(dirname "$0")/../gams_b200/csrc"
name=$1; ref=${2:-WORK}; shift; shift || true
mkdir -p build/variants /tmp/variant_$name
if [ "$ref" = WORK ]; then cp gc_wave_fused.cu /tmp/variant_$name/gc_wave_fused.cu; else git show $ref:gams_b200/csrc/gc_wave_fused.cu > /tmp/variant_$name/gc_wave_fused.cu; fi
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden -fmad=false -I. "$@" -c /tmp/variant_$name/gc_wave_fused.cu -o /tmp/variant_$name/fw.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/lib_$name.so build/api.o build/gc_window.o /tmp/variant_$name/fw.o build/zscore.o build/overlap.o build/ranges.o build/peaks.o build/gen.o build/idx_build.o build/count_windows.o build/inflate.o build/parse.o build/host.o -Xlinker --exclude-libs,ALL
echo built build/variants/lib_$name.so
