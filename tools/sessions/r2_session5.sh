#!/bin/bash
out=gpurun_out/r02_s5; mkdir -p $out
timeout 1700 python -m pytest tests/test_gpu_forward.py tests/test_gpu_pipeline.py -m gpu -q -k "full_row_groups or full_size or golden_tree or config1" 2>&1 | grep -v "^  *[a-z_]* = \|^$" | tail -150 > $out/pytest_fail.txt
