#!/bin/bash
# LSTM input projection: bias slice cached in shared memory (default) vs LDG per 16 columns (CB_TC_PROJ_VEC=0)
out=gpurun_out/r02_s32; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_forward.py tests/test_gpu_rna.py -x -q -m gpu > $out/pytest_forward.txt 2>&1; tail -2 $out/pytest_forward.txt
run() { name=$1; shift; echo "-- $name" >> $out/timing.txt
  env "$@" CB_PROF_DUMP=1 timeout 120 python tools/gpu_quick.py tc 4096 512 2>&1 | tail -3 >> $out/timing.txt; }
for rep in 1 2; do run vec_cached; run ldg CB_TC_PROJ_VEC=0; done
for cfg in "CB_TC_PROJ_VEC=1" "CB_TC_PROJ_VEC=0"; do echo "-- $cfg" >> $out/parity.txt; env $cfg timeout 300 python tools/parity_probe.py 32 >> $out/parity.txt 2>&1; done
cat $out/timing.txt; cut -c1-330 $out/parity.txt
