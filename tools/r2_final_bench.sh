#!/bin/bash
# final bench lines on the final library (after the evidence pass r02_e3, whose head kernel carried a reverted experiment)
out=gpurun_out/${1:-r02_final}; mkdir -p $out
python tools/box_speed.py > $out/box_speed.txt 2>&1; cat $out/box_speed.txt
timeout 1500 python -m pytest tests -m gpu -q > $out/pytest_gpu.txt 2>&1; tail -2 $out/pytest_gpu.txt
timeout 200 python __graft_entry__.py smoke > $out/smoke.txt 2>&1; tail -2 $out/smoke.txt
timeout 600 python bench.py > $out/bench_tc.json 2> $out/bench_tc.err; cut -c1-700 $out/bench_tc.json
timeout 400 python bench.py --impl reference > $out/bench_reference.json 2> $out/bench_reference.err; cut -c1-200 $out/bench_reference.json
for c in 2 3; do timeout 400 python bench.py --config $c --no-cpu-baseline > $out/bench_config$c.json 2> $out/bench_config$c.err; done
timeout 400 python bench.py --config 1 --steps 5 --warmup 1 > $out/bench_config1.json 2> $out/bench_config1.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 40 --csv --log-file $out/launches.csv \
    python tools/gpu_quick.py tc 4096 512 > $out/launches.log 2>&1
