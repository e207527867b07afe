#!/bin/bash
# source-level stall profiles of the recurrence, the input projection and the head
out=gpurun_out/r02_s31; mkdir -p $out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"lstm_tc_kernel" -s 1 -c 1 -f -o $out/prof_lstm \
    python tools/gpu_quick.py tc 4096 512 > $out/ncu_lstm.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel" -s 1 -c 1 -f -o $out/prof_proj \
    python tools/gpu_quick.py tc 4096 512 > $out/ncu_proj.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"head_tmajor" -c 1 -f -o $out/prof_head \
    python tools/gpu_quick.py tc 4096 512 > $out/ncu_head.log 2>&1
for k in lstm proj head; do ncu -i $out/prof_$k.ncu-rep --page source --csv --print-source sass > $out/${k}_source_sass.csv 2>/dev/null; gzip -f $out/${k}_source_sass.csv; done
rm -f $out/*.ncu-rep; ls -la $out
