#!/bin/bash
# recurrence: own half of K first (default) vs one sweep after the whole h arrived (CB_LSTM_SPLITK=0)
out=gpurun_out/r02_s21; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_forward.py -x -q -m gpu > $out/pytest_forward.txt 2>&1; tail -3 $out/pytest_forward.txt
run() { name=$1; shift; echo "-- $name" >> $out/timing.txt
  env "$@" CB_PROF_DUMP=1 timeout 120 python tools/gpu_quick.py tc 4096 512 2>&1 | tail -3 >> $out/timing.txt; }
for rep in 1 2; do run splitk; run sweep CB_LSTM_SPLITK=0; done
echo "-- B=1024 splitk / sweep" >> $out/timing.txt
CB_PROF_DUMP=1 timeout 120 python tools/gpu_quick.py tc 1024 512 2>&1 | tail -1 >> $out/timing.txt
CB_LSTM_SPLITK=0 CB_PROF_DUMP=1 timeout 120 python tools/gpu_quick.py tc 1024 512 2>&1 | tail -1 >> $out/timing.txt
for cfg in "CB_LSTM_SPLITK=1" "CB_LSTM_SPLITK=0"; do echo "-- $cfg" >> $out/parity.txt; env $cfg timeout 300 python tools/parity_probe.py 128 >> $out/parity.txt 2>&1; done
cat $out/timing.txt $out/parity.txt
