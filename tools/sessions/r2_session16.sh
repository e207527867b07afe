#!/bin/bash
# timeline probe of the partial-sum hand-over: K=256 conv2a (layer 4) and K=768 conv2b (layer 5), current defaults
out=gpurun_out/r02_s16; mkdir -p $out
for l in 4 5 6; do
  echo "== layer $l" >> $out/timeline.txt
  CB_TC_PROBE=$l CHIRON_B200_LIB=ab_libs/libTCDEV.so timeout 120 python tools/gpu_quick.py tc 4096 512 2>&1 | grep -A32 "^layer" | head -34 >> $out/timeline.txt
done
cat $out/timeline.txt
