#!/bin/bash
# A/B on one box: per-layer partial-sum length of the HBM-bound 1x1 convolutions, streaming stores in the epilogue
out=gpurun_out/r02_s15; mkdir -p $out
python tools/box_speed.py > $out/box.txt 2>&1
run() { # name, env...
  name=$1; shift
  echo "-- $name" >> $out/timing.txt
  env "$@" CB_PROF_DUMP=1 timeout 120 python tools/gpu_quick.py tc 4096 512 2>&1 | tail -3 >> $out/timing.txt
}
for rep in 1 2; do
run base
run k256_8 CB_TC_CPP_K256=8
run k256_8_k512_8 CB_TC_CPP_K256=8 CB_TC_CPP_K512=8
run stcs CHIRON_B200_LIB=ab_libs/libSTCS.so
run stcs_k256_8 CHIRON_B200_LIB=ab_libs/libSTCS.so CB_TC_CPP_K256=8
run all8 CB_TC_CPP=8
done
for cfg in "" "CB_TC_CPP_K256=8" "CB_TC_CPP_K256=8 CB_TC_CPP_K512=8"; do
  echo "-- $cfg" >> $out/parity.txt
  env $cfg timeout 300 python tools/parity_probe.py 192 >> $out/parity.txt 2>&1
done
cat $out/timing.txt $out/parity.txt
