#!/bin/bash
# usage: build_variant_of.sh NAME FILE.cu [extra nvcc flags...]: builds gams_b200/csrc/build/variants/lib_NAME.so with
# FILE.cu (working tree) recompiled with the extra flags, everything else from build/*.o (GAMS_CUDA_LIB=... selects it).
set -e
cd "$(dirname "$0")/../gams_b200/csrc"
name=$1; file=$2; shift; shift
base=${file%.cu}
mkdir -p build/variants /tmp/variant_$name
fm=""; case $base in zscore|gc_wave_fused|peaks|ranges) fm="-fmad=false";; esac
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden $fm -I. "$@" -c $file -o /tmp/variant_$name/v.o
objs=""; for o in api gc_window gc_wave_fused zscore overlap overlap_part ranges peaks gen idx_build count_windows inflate inflate2 parse host; do if [ $o = $base ]; then objs="$objs /tmp/variant_$name/v.o"; else objs="$objs build/$o.o"; fi; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/lib_$name.so $objs -Xlinker --exclude-libs,ALL
echo built build/variants/lib_$name.so
