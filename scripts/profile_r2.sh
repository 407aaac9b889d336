#!/bin/bash
# Run under gpurun (one GPU): round-2 ncu captures for profiles/.  Numbers printed by runs under ncu are never bench values.
set -u
OUT=gpurun_out
R=${1:-r2}
KERNELS="lut_q8_kernel|q8_beam_kernel|q8_search_kernel|rerank_kernel|exact_tc_kernel|tc_select_kernel|tc_threshold_kernel|tc_convert|exact_kernel|encode_kernel|merge_kernel|search_kernel"
# (1) every launch of OUR kernels in a short bench run with its device time (cold, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$KERNELS" -c 3000 --csv --log-file $OUT/launches_${R}.csv \
    python bench.py --steps 3 --warmup 3 --quiet --no-shards > /dev/null 2>&1
# (2) full captures: the traversal (first launch after warm-up), the table build, the rerank, the tensor-core brute force (pass B)
ncu --set full --clock-control none --import-source on -k regex:q8_beam -s 3 -c 1 -o $OUT/prof_${R}_beam python bench.py --steps 1 --warmup 3 --quiet --no-shards > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:lut_q8 -s 3 -c 1 -o $OUT/prof_${R}_lut python bench.py --steps 1 --warmup 3 --quiet --no-shards > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:rerank_kernel -s 3 -c 1 -o $OUT/prof_${R}_rerank python bench.py --steps 1 --warmup 3 --quiet --no-shards > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:exact_tc_kernelILi1 -c 1 -o $OUT/prof_${R}_exact_tc python bench.py --steps 1 --warmup 3 --quiet --no-shards > /dev/null 2>&1
ls -la $OUT/*${R}*
