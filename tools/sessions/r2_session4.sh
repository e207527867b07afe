#!/bin/bash
# Round-2 GPU session 4: GPU test suite on the parity defaults, smoke, three-pass beam search on real logits, bench lines.
out=gpurun_out/r02_s4; mkdir -p $out
timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -30 | tee $out/pytest_gpu.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -4 | tee $out/smoke.txt
echo "== beam search on real logits" | tee $out/beam_real_ab.jsonl
timeout 600 python tools/experiments/beam_real_ab.py 2>&1 | tail -8 | tee -a $out/beam_real_ab.jsonl
timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 > $out/bench_tc.json; cat $out/bench_tc.json | cut -c1-1500
timeout 300 python bench.py --config 2 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > $out/bench_config2.json
timeout 300 python bench.py --config 3 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > $out/bench_config3.json
timeout 300 python bench.py --config 1 --steps 3 --warmup 1 2>&1 | tail -1 > $out/bench_config1.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > $out/bench_reference.json
for f in config2 config3 config1 reference; do cut -c1-700 $out/bench_$f.json; done
