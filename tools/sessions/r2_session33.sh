#!/bin/bash
out=gpurun_out/r02_s33; mkdir -p $out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel" -s 2 -c 1 -f -o $out/prof_proj \
    python tools/gpu_quick.py tc 4096 512 > $out/ncu_proj.log 2>&1
ncu -i $out/prof_proj.ncu-rep --page source --csv --print-source sass > $out/proj_source_sass.csv 2>/dev/null; gzip -f $out/proj_source_sass.csv
ncu -i $out/prof_proj.ncu-rep --page raw --csv > $out/proj_raw.csv 2>/dev/null
rm -f $out/*.ncu-rep; ls -la $out
