#!/bin/bash
# Session-3 GPU pass (one short gpurun call): all GPU tests, smoke, a bench line, timing + ncu traffic of the bn_* kernels,
# and the file-level `chiron call` throughput (tools/call_bench.py).
out=gpurun_out/${1:-r01_s3}
mkdir -p $out
timeout 200 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.txt 2>&1; tail -15 $out/pytest_gpu.txt
timeout 60 python tools/bn_bench.py 1024 512 3 > $out/bn_bench.json 2> $out/bn_bench.err; cat $out/bn_bench.json
timeout 150 python bench.py --steps 10 --warmup 3 > $out/bench_tc.json 2> $out/bench_tc.err; cut -c1-300 $out/bench_tc.json
timeout 90 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:bn_ --csv --log-file $out/bn_kernels.csv python tools/bn_bench.py 1024 512 1 > $out/bn_ncu.log 2>&1
tail -3 $out/bn_kernels.csv
timeout 60 python tools/call_bench.py --reads 600 --fmt signal > $out/call_signal.json 2> $out/call_signal.err; tail -1 $out/call_signal.json
timeout 60 python tools/call_bench.py --reads 600 --fmt fast5 > $out/call_fast5.json 2> $out/call_fast5.err; tail -1 $out/call_fast5.json
timeout 60 python tools/call_bench.py --reads 600 --fmt signal --stub > $out/call_signal_stub.json 2> $out/call_signal_stub.err; tail -1 $out/call_signal_stub.json
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $out/smoke.txt 2>&1; tail -1 $out/smoke.txt
