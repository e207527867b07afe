#!/bin/bash
# Round-2 GPU session 3: calibration of the truncation-compensation constant (CB_TC_BIAS) at 1 / 2-3 partial sums per tile.
out=gpurun_out/r02_s3; mkdir -p $out
echo "== diag (float64 oracle, 96 read1 windows)" | tee $out/diag.txt
for cpp in 0 12 8; do for bias in 0.25 0.3 0.35 0.4 0.45; do echo "-- CB_TC_CPP=$cpp CB_TC_BIAS=$bias" | tee -a $out/diag.txt; CB_TC_CPP=$cpp CB_TC_BIAS=$bias timeout 300 python tools/gpu_diag.py 96 tc 2>&1 | tail -2 | tee -a $out/diag.txt; done; done
echo "== parity on the bench batch" | tee $out/parity.txt
for cpp in 0 12 8; do for bias in 0.3 0.35 0.4; do CB_TC_CPP=$cpp CB_TC_BIAS=$bias timeout 600 python tools/parity_probe.py 64 2>&1 | tail -1 | tee -a $out/parity.txt; done; done
echo "== timing" | tee $out/timing.txt
for cpp in 0 12 8; do echo "-- CB_TC_CPP=$cpp" | tee -a $out/timing.txt; CB_TC_CPP=$cpp timeout 200 python tools/gpu_quick.py tc 4096 512 2>&1 | tail -1 | tee -a $out/timing.txt; done
