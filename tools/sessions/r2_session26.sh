#!/bin/bash
# source-level stall profile of the beam search on real logits (pass 1 and the retry pass)
out=gpurun_out/r02_s28; mkdir -p $out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"beam_warp|beam_retry" -c 2 -f -o $out/prof_beam \
    python tools/experiments/beam_one.py > $out/ncu_beam.log 2>&1
ncu -i $out/prof_beam.ncu-rep --page source --csv --print-source sass > $out/beam_source_sass.csv 2>/dev/null
ncu -i $out/prof_beam.ncu-rep --page source --csv --print-source cuda > $out/beam_source_cuda.csv 2>/dev/null
gzip -f $out/beam_source_sass.csv $out/beam_source_cuda.csv
rm -f $out/prof_beam.ncu-rep; ls -la $out
