#!/bin/bash
# lean conv epilogue (EPI template) vs the generic one: same-box timing, identical numerics, hand-over timeline
out=gpurun_out/r02_s17; mkdir -p $out
run() { name=$1; shift; echo "-- $name" >> $out/timing.txt
  env "$@" CB_PROF_DUMP=1 timeout 120 python tools/gpu_quick.py tc 4096 512 2>&1 | tail -3 >> $out/timing.txt; }
for rep in 1 2; do run lean; run generic CB_TC_LEAN_EPI=0; done
run lean_b1024 ; 
for cfg in "CB_TC_LEAN_EPI=1" "CB_TC_LEAN_EPI=0"; do echo "-- $cfg" >> $out/parity.txt; env $cfg timeout 300 python tools/parity_probe.py 64 >> $out/parity.txt 2>&1; done
for l in 4 5 6; do
  echo "== layer $l" >> $out/timeline.txt
  CB_TC_PROBE=$l CHIRON_B200_LIB=ab_libs/libTCDEV.so timeout 120 python tools/gpu_quick.py tc 4096 512 2>&1 | grep -A14 "^layer" | head -15 >> $out/timeline.txt
done
cat $out/timing.txt $out/parity.txt $out/timeline.txt
