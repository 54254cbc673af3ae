#!/bin/bash
# usage (under gpurun): bash scripts/r2_variants.sh  -> count_probe.py for the default library and every build/variants/*.so
for f in default gams_b200/csrc/build/variants/*.so; do [ "$f" = default ] || [ -e "$f" ] || continue
if [ "$f" = default ]; then timeout 200 python scripts/count_probe.py 2>&1 | tail -1; else GAMS_CUDA_LIB=$PWD/$f timeout 200 python scripts/count_probe.py 2>&1 | tail -1; fi
done
