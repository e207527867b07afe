#!/bin/bash
# Round-2 GPU session 1: numerics of the partial-sum tensor-core path (CB_TC_CPP sweep) against the float64 oracle and
# the fp32 mode, timing A/B against the round-1 library (ab_libs/libA.so), LSTM timeline probe, GPU test suite.
out=gpurun_out/r02_s1; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/gpu.txt
echo "== diag (float64 oracle, 96 read1 windows)" | tee $out/diag.txt
timeout 300 python tools/gpu_diag.py 96 fp32 2>&1 | tee -a $out/diag.txt
for cpp in 1 2 4 64; do echo "-- CB_TC_CPP=$cpp" | tee -a $out/diag.txt; CB_TC_CPP=$cpp timeout 300 python tools/gpu_diag.py 96 tc 2>&1 | tail -1 | tee -a $out/diag.txt; done
echo "-- round-1 library" | tee -a $out/diag.txt
CHIRON_B200_LIB=ab_libs/libA.so timeout 300 python tools/gpu_diag.py 96 tc 2>&1 | tail -1 | tee -a $out/diag.txt
echo "== timing 4096x512" | tee $out/timing.txt
for rep in 1 2; do
  echo "-- round-1 library" | tee -a $out/timing.txt; CHIRON_B200_LIB=ab_libs/libA.so timeout 200 python tools/gpu_quick.py tc 4096 512 2>&1 | tail -1 | tee -a $out/timing.txt
  for cpp in 1 2 4 64; do echo "-- CB_TC_CPP=$cpp" | tee -a $out/timing.txt; CB_TC_CPP=$cpp timeout 200 python tools/gpu_quick.py tc 4096 512 2>&1 | tail -1 | tee -a $out/timing.txt; done
done
echo "-- B=1024" | tee -a $out/timing.txt; timeout 200 python tools/gpu_quick.py tc 1024 512 2>&1 | tail -1 | tee -a $out/timing.txt
echo "== parity on the bench batch" | tee $out/parity.txt
for cpp in 1 2 4; do CB_TC_CPP=$cpp timeout 600 python tools/parity_probe.py 64 2>&1 | tail -1 | tee -a $out/parity.txt; done
CHIRON_B200_LIB=ab_libs/libA.so timeout 600 python tools/parity_probe.py 0 2>&1 | tail -1 | tee -a $out/parity.txt
echo "== lstm timeline probe" | tee $out/lstm_probe.txt
CB_LSTM_PROBE=1 CHIRON_B200_LIB=ab_libs/libDEV.so timeout 200 python tools/gpu_quick.py tc 4096 512 2>&1 | tail -120 >> $out/lstm_probe.txt
echo "== pytest" 
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee $out/pytest_gpu.txt
echo "== beam search on real logits" | tee gpurun_out/r02_s1/beam_real_ab.jsonl
timeout 400 python tools/experiments/beam_real_ab.py 2>&1 | tail -8 | tee -a gpurun_out/r02_s1/beam_real_ab.jsonl
