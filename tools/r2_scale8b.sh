#!/bin/bash
# files -> fastq on 1, 4, 8 GPUs of one box with the final pipeline (cb_reserve_sms, beam filters), greedy and beam 30
out=gpurun_out/r02_scale_b; mkdir -p $out
python tools/call_scale.py --prepare --reads 6400 --fmt signal --dir /dev/shm/cs_signal > $out/prepare.txt 2>&1
for n in 1 4 8; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) \
    tools/call_scale.py --dir /dev/shm/cs_signal --out /dev/shm/cs_out > $out/call_n$n.json 2> $out/call_n$n.err
  tail -1 $out/call_n$n.json | cut -c1-330
done
for n in 1 8; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29530+n)) \
    tools/call_scale.py --dir /dev/shm/cs_signal --out /dev/shm/cs_out --beam 30 > $out/call_n${n}_beam30.json 2> $out/call_n${n}_beam30.err
tail -1 $out/call_n${n}_beam30.json | cut -c1-330
done
