#!/bin/bash
# cb_reserve_sms in evaluation(): files -> fastq with / without, greedy and beam 30, fast5; pipeline tests
out=gpurun_out/r02_s24; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_pipeline.py -x -q -m gpu > $out/pytest_pipeline.txt 2>&1; tail -2 $out/pytest_pipeline.txt
for rep in 1 2; do
for r in 4 0; do
echo "-- CHIRON_B200_RESERVE_SMS=$r greedy" >> $out/call.txt
CHIRON_B200_RESERVE_SMS=$r timeout 120 python tools/call_bench.py --reads 1600 --fmt signal 2>/dev/null | tail -1 >> $out/call.txt
done; done
for r in 4 0; do
echo "-- CHIRON_B200_RESERVE_SMS=$r beam 30" >> $out/call.txt
CHIRON_B200_RESERVE_SMS=$r timeout 120 python tools/call_bench.py --reads 1600 --fmt signal --beam 30 2>/dev/null | tail -1 >> $out/call.txt
echo "-- CHIRON_B200_RESERVE_SMS=$r fast5" >> $out/call.txt
CHIRON_B200_RESERVE_SMS=$r timeout 120 python tools/call_bench.py --reads 1600 --fmt fast5 2>/dev/null | tail -1 >> $out/call.txt
done
cat $out/call.txt
