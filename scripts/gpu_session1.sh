#!/bin/bash
# round 2, GPU session 1: first run of the pipelined traversal kernel (tests, sanitizer, sweep against the synchronous kernel)
set -u
OUT=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $OUT/s1_smi.log 2>&1
timeout 300 python scripts/sanitize_q8.py 1500 > $OUT/s1_plain.log 2>&1; echo "plain rc=$?" >> $OUT/s1_plain.log
timeout 600 compute-sanitizer --tool memcheck python scripts/sanitize_q8.py 1500 > $OUT/s1_memcheck.log 2>&1; echo "rc=$?" >> $OUT/s1_memcheck.log
timeout 900 python -m pytest tests/test_gpu_q8.py -q -x --timeout 300 > $OUT/s1_tests.log 2>&1; echo "rc=$?" >> $OUT/s1_tests.log
timeout 900 python scripts/k2_sweep.py > $OUT/s1_sweep.log 2> $OUT/s1_sweep.err; echo "rc=$?" >> $OUT/s1_sweep.err
tail -3 $OUT/s1_plain.log; tail -3 $OUT/s1_memcheck.log; tail -5 $OUT/s1_tests.log; cat $OUT/s1_sweep.log
