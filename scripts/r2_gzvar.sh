#!/bin/bash
# usage (under gpurun): bash scripts/r2_gzvar.sh -> gz_probe.py (decoder timings) for the default library and every build/variants/*.so
for f in default gams_b200/csrc/build/variants/*.so; do [ "$f" = default ] || [ -e "$f" ] || continue
echo "== $f"
if [ "$f" = default ]; then timeout 300 python scripts/gz_probe.py 2>&1 | tail -3 | head -2; else GAMS_CUDA_LIB=$PWD/$f timeout 300 python scripts/gz_probe.py 2>&1 | tail -3 | head -2; fi
done
