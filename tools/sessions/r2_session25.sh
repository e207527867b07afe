#!/bin/bash
# compute-sanitizer over the round-2 kernels (tc forward with partial sums + lean epilogue, three-pass beam with scores,
# assembly, submit/collect) on small shapes: memcheck, then racecheck and synccheck on the shared-memory / mbarrier code
out=gpurun_out/r02_s25; mkdir -p $out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/experiments/sanitize_small.py > $out/memcheck.txt 2>&1; echo "memcheck rc=$?" >> $out/memcheck.txt
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 9 python tools/experiments/sanitize_small.py > $out/synccheck.txt 2>&1; echo "synccheck rc=$?" >> $out/synccheck.txt
tail -n 6 $out/memcheck.txt; tail -n 6 $out/synccheck.txt
