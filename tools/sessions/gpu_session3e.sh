#!/bin/bash
# Session-3e GPU pass (the round's last seconds of GPU budget): beam-search tests first, then the staged / global-prefetch A/B of
# beam_warp_kernel, then the rest of the GPU tests.
out=gpurun_out/${1:-r01_s3e}
mkdir -p $out
timeout 40 python -m pytest tests -m gpu -x -q -k "beam or rna or golden_tree or decode" > $out/pytest_beam.txt 2>&1; tail -2 $out/pytest_beam.txt
timeout 40 python tools/experiments/beam_stage_ab.py > $out/beam_stage_ab.jsonl 2> $out/beam_stage_ab.err; cat $out/beam_stage_ab.jsonl
timeout 60 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.txt 2>&1; tail -2 $out/pytest_gpu.txt
