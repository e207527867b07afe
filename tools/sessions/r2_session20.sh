#!/bin/bash
# 7-stage operand ring for the plain convolutions vs 6; stagger sweep with the lean epilogue
out=gpurun_out/r02_s20; mkdir -p $out
run() { name=$1; shift; echo "-- $name" >> $out/timing.txt
  env "$@" CB_PROF_DUMP=1 timeout 120 python tools/gpu_quick.py tc 4096 512 2>&1 | tail -3 >> $out/timing.txt; }
for rep in 1 2; do run stages7; run stages6 CB_TC_STAGES=6; done
run stagger0 CB_TC_STAGGER=0; run stagger2 CB_TC_STAGGER=2; run stagger8 CB_TC_STAGGER=8
run cpp3 CB_TC_CPP=3
timeout 300 python tools/parity_probe.py 32 >> $out/parity.txt 2>&1
cat $out/timing.txt $out/parity.txt
