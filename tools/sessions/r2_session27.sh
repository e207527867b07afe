#!/bin/bash
# beam search: idle survivors skipped while nothing was evicted + shift indexing -- GPU tests, then timing on real logits
out=gpurun_out/r02_s29; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_decode_assemble.py tests/test_gpu_rna.py -x -q -m gpu > $out/pytest_decode.txt 2>&1; tail -3 $out/pytest_decode.txt
timeout 600 python tools/experiments/beam_real_ab.py > $out/beam_real_ab.jsonl 2> $out/beam_real_ab.err
cut -c1-420 $out/beam_real_ab.jsonl; tail -3 $out/beam_real_ab.err
timeout 120 python tools/call_bench.py --reads 1600 --fmt signal --beam 30 2>/dev/null | tail -1 > $out/call_beam30.json; cat $out/call_beam30.json
