#!/bin/bash
# h exchange by remote stores (default) vs bulk DSMEM copies (CB_LSTM_EXCH=bulk): parity first, then same-box timing
out=gpurun_out/r02_s19; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_forward.py -x -q -m gpu > $out/pytest_forward.txt 2>&1; tail -5 $out/pytest_forward.txt
run() { name=$1; shift; echo "-- $name" >> $out/timing.txt
  env "$@" CB_PROF_DUMP=1 timeout 120 python tools/gpu_quick.py tc 4096 512 2>&1 | tail -3 >> $out/timing.txt; }
for rep in 1 2; do run remote; run bulk CB_LSTM_EXCH=bulk; done
echo "-- B=1024" >> $out/timing.txt
CB_PROF_DUMP=1 timeout 120 python tools/gpu_quick.py tc 1024 512 2>&1 | tail -2 >> $out/timing.txt
CB_LSTM_EXCH=bulk CB_PROF_DUMP=1 timeout 120 python tools/gpu_quick.py tc 1024 512 2>&1 | tail -2 >> $out/timing.txt
for cfg in "CB_LSTM_EXCH=remote" "CB_LSTM_EXCH=bulk"; do echo "-- $cfg" >> $out/parity.txt; env $cfg timeout 300 python tools/parity_probe.py 64 >> $out/parity.txt 2>&1; done
cat $out/timing.txt $out/parity.txt
