#!/bin/bash
# Round evidence pass (run under gpurun on one B200): GPU tests, bench arms, ncu launch list and full captures.
# Usage: tools/gpu_evidence.sh <tag>   -> gpurun_out/<tag>/   (summarise here with tools/ncu_summary.py / ncu_traffic.py)
tag=${1:-r01}
out=gpurun_out/$tag
mkdir -p $out
python tools/box_speed.py > $out/box_speed.txt 2>&1; cat $out/box_speed.txt
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.txt 2>&1; tail -2 $out/pytest_gpu.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $out/smoke.txt 2>&1; tail -1 $out/smoke.txt
timeout 400 python bench.py > $out/bench_tc.json 2> $out/bench_tc.err; cat $out/bench_tc.json | cut -c1-400
timeout 400 python bench.py --precision fp32 --no-cpu-baseline --steps 5 > $out/bench_fp32.json 2> $out/bench_fp32.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_reference.json 2> $out/bench_reference.err
timeout 600 python tools/bench_configs.py tc > $out/configs_tc.json 2> $out/configs_tc.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 11 -c 11 -f -o $out/prof_gemm \
    python tools/gpu_quick.py tc 4096 512 > $out/ncu_gemm.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_tc_kernel -s 3 -c 3 -f -o $out/prof_lstm \
    python tools/gpu_quick.py tc 4096 512 > $out/ncu_lstm.log 2>&1
# batch-statistics BatchNorm kernels: fp32 forward in both BN modes + ncu launch list with DRAM bytes (tools/ncu_bn_summary.py)
timeout 60 python tools/bn_bench.py 1024 512 3 > $out/bn_bench.json 2> $out/bn_bench.err
timeout 120 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:bn_ --csv --log-file $out/bn_kernels.csv python tools/bn_bench.py 1024 512 1 > $out/bn_ncu.log 2>&1
# `chiron call` on files -> fastq (and the host pipeline alone), beam-search variants
timeout 90 python tools/call_bench.py --reads 800 --fmt signal > $out/call_signal.json 2> $out/call_signal.err
timeout 90 python tools/call_bench.py --reads 800 --fmt fast5 > $out/call_fast5.json 2> $out/call_fast5.err
timeout 90 python tools/call_bench.py --reads 800 --fmt signal --beam 30 > $out/call_signal_beam30.json 2> $out/call_signal_beam30.err
timeout 90 python tools/call_bench.py --reads 800 --fmt signal --stub > $out/call_signal_stub.json 2> $out/call_signal_stub.err
timeout 120 python tools/experiments/beam_stage_ab.py > $out/beam_stage_ab.jsonl 2> $out/beam_stage_ab.err
timeout 300 python tools/experiments/beam_real_ab.py > $out/beam_real_ab.jsonl 2> $out/beam_real_ab.err     # real logits: the numbers that count
ls -la $out
