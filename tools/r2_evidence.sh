#!/bin/bash
# Round-2 evidence pass (run under gpurun on ONE B200): GPU tests, smoke, bench arms and configs, ncu launch list and full
# captures of every kernel class on the path, beam search on real logits, files -> fastq on one GPU.
# Usage: tools/r2_evidence.sh <tag>   -> gpurun_out/<tag>/   (summarise here with tools/ncu_summary.py / ncu_traffic.py)
tag=${1:-r02_final}
out=gpurun_out/$tag
mkdir -p $out
python tools/box_speed.py > $out/box_speed.txt 2>&1; cat $out/box_speed.txt
timeout 1500 python -m pytest tests -m gpu -q > $out/pytest_gpu.txt 2>&1; tail -2 $out/pytest_gpu.txt
timeout 200 python __graft_entry__.py smoke > $out/smoke.txt 2>&1; tail -2 $out/smoke.txt
timeout 600 python bench.py > $out/bench_tc.json 2> $out/bench_tc.err; cut -c1-300 $out/bench_tc.json
timeout 400 python bench.py --impl reference > $out/bench_reference.json 2> $out/bench_reference.err; cut -c1-200 $out/bench_reference.json
timeout 400 python bench.py --precision fp32 --no-cpu-baseline --no-parity --steps 5 > $out/bench_fp32.json 2> $out/bench_fp32.err
for c in 2 3; do timeout 400 python bench.py --config $c --no-cpu-baseline > $out/bench_config$c.json 2> $out/bench_config$c.err; done
timeout 400 python bench.py --config 1 --steps 5 --warmup 1 > $out/bench_config1.json 2> $out/bench_config1.err
timeout 600 python tools/bench_configs.py tc > $out/configs_tc.json 2> $out/configs_tc.err
# launch list of one forward + decode (cold-cache, serialised: shares, not absolutes)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 40 --csv --log-file $out/launches.csv \
    python tools/gpu_quick.py tc 4096 512 > $out/launches.log 2>&1
# full captures: the conv contractions (CTA-pair kernel: K=768 with partial sums, K=256), the input projection, the recurrence,
# head / generator / transpose / greedy / path_prob, beam search, assembly
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 11 -c 11 -f -o $out/prof_gemm \
    python tools/gpu_quick.py tc 4096 512 > $out/ncu_gemm.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:lstm_tc_kernel -s 3 -c 3 -f -o $out/prof_lstm \
    python tools/gpu_quick.py tc 4096 512 > $out/ncu_lstm.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"head_tmajor|gen_conv2a|transpose_x|greedy|path_prob|seq_len" -s 6 -c 6 -f -o $out/prof_small \
    python tools/gpu_quick.py tc 4096 512 > $out/ncu_small.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"beam_" -c 3 -f -o $out/prof_beam \
    python tools/experiments/beam_one.py > $out/ncu_beam.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"asm_" -c 10 -f -o $out/prof_asm \
    python tools/experiments/asm_one.py > $out/ncu_asm.log 2>&1
# `chiron call` on files -> fastq (and the host pipeline alone), beam search on real logits
timeout 120 python tools/call_bench.py --reads 800 --fmt signal > $out/call_signal.json 2> $out/call_signal.err
timeout 120 python tools/call_bench.py --reads 800 --fmt fast5 > $out/call_fast5.json 2> $out/call_fast5.err
timeout 120 python tools/call_bench.py --reads 800 --fmt signal --beam 30 > $out/call_signal_beam30.json 2> $out/call_signal_beam30.err
timeout 120 python tools/call_bench.py --reads 800 --fmt signal --stub > $out/call_signal_stub.json 2> $out/call_signal_stub.err
timeout 400 python tools/experiments/beam_real_ab.py > $out/beam_real_ab.jsonl 2> $out/beam_real_ab.err
cat $out/call_signal.json $out/call_fast5.json $out/call_signal_beam30.json
# gpurun brings back at most 64 MiB: summarise the captures on the box and drop the reports
for r in gemm lstm small beam asm; do python tools/ncu_summary.py $out/prof_$r.ncu-rep > $out/ncu_$r.txt 2>&1; done
python tools/ncu_traffic.py $out tc 4096 > $out/traffic.json 2> $out/traffic.err
ncu -i $out/prof_gemm.ncu-rep --page source --csv > $out/ncu_gemm_source.csv 2>/dev/null; gzip -f $out/ncu_gemm_source.csv
rm -f $out/*.ncu-rep
ls -la $out; du -sh $out
