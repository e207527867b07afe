#!/bin/bash
# do a few SMs left free by the persistent GEMM grids pay for themselves in `chiron call` (assembly kernels find a slot)?
out=gpurun_out/r02_s23; mkdir -p $out
for rep in 1 2; do
for sms in 148 146 144 140; do
echo "-- CB_TC_GEMM_SMS=$sms" >> $out/call.txt
CB_TC_GEMM_SMS=$sms timeout 120 python tools/call_bench.py --reads 1600 --fmt signal 2>/dev/null | tail -1 >> $out/call.txt
done
done
echo "-- CB_TC_GEMM_SMS=146 resident step" >> $out/call.txt
CB_TC_GEMM_SMS=146 timeout 120 python tools/gpu_quick.py tc 4096 400 2>&1 | tail -1 >> $out/call.txt
timeout 120 python tools/gpu_quick.py tc 4096 400 2>&1 | tail -1 >> $out/call.txt
cat $out/call.txt
