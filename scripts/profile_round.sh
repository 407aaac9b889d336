#!/bin/bash
# Run under gpurun (one GPU): ncu captures for profiles/.  Numbers printed by runs under ncu are never bench values.
set -u
OUT=gpurun_out
R=${1:-r1}
# (1) every launch of OUR kernels in a short bench run with its device time (cold, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"lut_q8_kernel|q8_search_kernel|rerank_kernel|exact_kernel|encode_kernel|merge_kernel|search_kernel" -c 3000 --csv --log-file $OUT/launches_${R}.csv \
    python bench.py --steps 3 --warmup 3 --quiet > /dev/null 2>&1
# (2) full captures of the dominant kernels (first launch after warm-up)
ncu --set full --clock-control none --import-source on -k regex:fast_search -s 3 -c 1 -o $OUT/prof_${R}_search python bench.py --steps 1 --warmup 3 --quiet > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:rerank_kernel -s 3 -c 1 -o $OUT/prof_${R}_rerank python bench.py --steps 1 --warmup 3 --quiet > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:encode_kernel -c 1 -o $OUT/prof_${R}_encode python bench.py --steps 1 --warmup 3 --quiet > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:exact_kernel -s 1 -c 1 -o $OUT/prof_${R}_exact python bench.py --steps 1 --warmup 3 --quiet > /dev/null 2>&1
ls -la $OUT/*.ncu-rep
