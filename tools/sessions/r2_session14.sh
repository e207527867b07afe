#!/bin/bash
out=gpurun_out/r02_s14; mkdir -p $out
timeout 300 python tools/gpu_diag.py 96 tc fp32 2>&1 | tail -3 | tee $out/diag.txt
timeout 600 python tools/parity_probe.py 512 2>&1 | tail -1 | tee $out/parity.txt
timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee $out/pytest_gpu.txt
timeout 600 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 > $out/bench_tc.json; cut -c1-600 $out/bench_tc.json
