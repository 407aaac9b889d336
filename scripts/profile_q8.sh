#!/bin/bash
# Run under gpurun (one GPU): ncu captures of the 8-bit table path.  Numbers printed by runs under ncu are never bench values.
set -u
OUT=gpurun_out
R=${1:-r1q8}
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"lut_q8_kernel|q8_search_kernel|rerank_kernel|exact_kernel|encode_kernel|merge_kernel|search_kernel" -c 3000 --csv --log-file $OUT/launches_${R}.csv \
    python bench.py --steps 3 --warmup 3 --quiet > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:q8_search -s 3 -c 1 -o $OUT/prof_${R}_search python bench.py --steps 1 --warmup 3 --quiet > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:lut_q8 -s 3 -c 1 -o $OUT/prof_${R}_lut python bench.py --steps 1 --warmup 3 --quiet > /dev/null 2>&1
ls -la $OUT/*${R}*
