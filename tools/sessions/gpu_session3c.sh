#!/bin/bash
# Session-3c GPU pass: all GPU tests (GRU cells, refactored LSTM launcher, finisher thread) + A/B of the finisher thread.
out=gpurun_out/${1:-r01_s3c}
mkdir -p $out
timeout 240 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.txt 2>&1; tail -15 $out/pytest_gpu.txt
for i in 1 2; do
  for f in 1 0; do
    CHIRON_B200_FINISHER=$f timeout 60 python tools/call_bench.py --reads 800 --fmt signal 2>> $out/call.err | tail -1 | sed "s/^{/{\"finisher\": $f, /" | tee -a $out/call_ab.jsonl
  done
done
CHIRON_B200_FINISHER=1 timeout 60 python tools/call_bench.py --reads 800 --fmt fast5 2>> $out/call.err | tail -1 | sed "s/^{/{\"finisher\": 1, /" | tee -a $out/call_ab.jsonl
