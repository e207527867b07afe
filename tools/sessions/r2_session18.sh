#!/bin/bash
# recurrence timeline (gate warps of CTA (0,0), steps 100..103) + ablations
out=gpurun_out/r02_s18; mkdir -p $out
CB_LSTM_PROBE=1 CHIRON_B200_LIB=ab_libs/libDEV.so timeout 120 python tools/gpu_quick.py tc 4096 512 > $out/probe.txt 2>&1
for f in 0 1 2 4 7; do echo "-- dbg_flags $f" >> $out/ablate.txt; CB_LSTM_DBG=$f CB_PROF_DUMP=1 CHIRON_B200_LIB=ab_libs/libDEV.so timeout 120 python tools/gpu_quick.py tc 4096 512 2>&1 | tail -1 >> $out/ablate.txt; done
head -120 $out/probe.txt; cat $out/ablate.txt
