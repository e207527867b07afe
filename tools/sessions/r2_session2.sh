#!/bin/bash
# Round-2 GPU session 2: truncation compensation (CB_TC_BIAS sweep) with one accumulator per tile, partial sums of 8 chunks,
# timing against the round-1 library.
out=gpurun_out/r02_s2; mkdir -p $out
echo "== diag (float64 oracle, 96 read1 windows)" | tee $out/diag.txt
for bias in 0 0.5 0.7213 1.0 1.4; do echo "-- CB_TC_BIAS=$bias" | tee -a $out/diag.txt; CB_TC_BIAS=$bias timeout 300 python tools/gpu_diag.py 96 tc 2>&1 | tail -2 | tee -a $out/diag.txt; done
for cpp in 8 4; do echo "-- CB_TC_CPP=$cpp (default bias)" | tee -a $out/diag.txt; CB_TC_CPP=$cpp timeout 300 python tools/gpu_diag.py 96 tc 2>&1 | tail -2 | tee -a $out/diag.txt; done
echo "-- CB_TC_CPP=2 CB_TC_BIAS=0: accurate convolutions + round-1 recurrence" | tee -a $out/diag.txt; CB_TC_BIAS=0 CB_TC_CPP=2 timeout 300 python tools/gpu_diag.py 96 tc 2>&1 | tail -2 | tee -a $out/diag.txt
echo "-- CB_TC_CPP=2 (default bias)" | tee -a $out/diag.txt; CB_TC_CPP=2 timeout 300 python tools/gpu_diag.py 96 tc 2>&1 | tail -2 | tee -a $out/diag.txt
echo "== timing 4096x512" | tee $out/timing.txt
for rep in 1 2; do
  echo "-- round-1 library" | tee -a $out/timing.txt; CHIRON_B200_LIB=ab_libs/libA.so timeout 200 python tools/gpu_quick.py tc 4096 512 2>&1 | tail -1 | tee -a $out/timing.txt
  echo "-- default" | tee -a $out/timing.txt; timeout 200 python tools/gpu_quick.py tc 4096 512 2>&1 | tail -1 | tee -a $out/timing.txt
  for cpp in 8 4; do echo "-- CB_TC_CPP=$cpp" | tee -a $out/timing.txt; CB_TC_CPP=$cpp timeout 200 python tools/gpu_quick.py tc 4096 512 2>&1 | tail -1 | tee -a $out/timing.txt; done
done
echo "== parity on the bench batch" | tee $out/parity.txt
timeout 600 python tools/parity_probe.py 64 2>&1 | tail -1 | tee -a $out/parity.txt
CB_TC_CPP=8 timeout 600 python tools/parity_probe.py 64 2>&1 | tail -1 | tee -a $out/parity.txt
