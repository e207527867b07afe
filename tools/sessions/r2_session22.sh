#!/bin/bash
# what the per-read assembly launches cost `chiron call` (files -> fastq, 1 GPU)
out=gpurun_out/r02_s22; mkdir -p $out
for rep in 1 2; do
timeout 120 python tools/call_bench.py --reads 1600 --fmt signal 2>/dev/null | tail -1 >> $out/call.txt
CHIRON_B200_SKIP_ASSEMBLY=1 timeout 120 python tools/call_bench.py --reads 1600 --fmt signal 2>/dev/null | tail -1 >> $out/call.txt
done
CHIRON_B200_FINISHER=0 timeout 120 python tools/call_bench.py --reads 1600 --fmt signal 2>/dev/null | tail -1 >> $out/call.txt
timeout 120 python tools/gpu_quick.py tc 4096 400 2>&1 | tail -1 >> $out/call.txt
cat $out/call.txt
