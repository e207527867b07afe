#!/bin/bash
# files -> fastq on 1, 2, 4, 8 GPUs of one box (read-sharded chiron call), plus the host pipeline alone at N = 8
out=gpurun_out/r02_scale; mkdir -p $out
nproc > $out/host.txt; nvidia-smi -L >> $out/host.txt; free -g | head -2 >> $out/host.txt
python tools/call_scale.py --prepare --reads 6400 --fmt signal --dir /dev/shm/cs_signal > $out/prepare.txt 2>&1
for n in 1 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) \
    tools/call_scale.py --dir /dev/shm/cs_signal --out /dev/shm/cs_out > $out/call_n$n.json 2> $out/call_n$n.err
  tail -1 $out/call_n$n.json | cut -c1-400
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 \
    tools/call_scale.py --dir /dev/shm/cs_signal --out /dev/shm/cs_out --stub > $out/call_n8_stub.json 2> $out/call_n8_stub.err
tail -1 $out/call_n8_stub.json | cut -c1-400
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29530 \
    tools/call_scale.py --dir /dev/shm/cs_signal --out /dev/shm/cs_out --beam 30 > $out/call_n8_beam30.json 2> $out/call_n8_beam30.err
tail -1 $out/call_n8_beam30.json | cut -c1-400
for f in $out/*.err; do grep -v "^Found\|^$\|OMP_NUM_THREADS\|\*\*\*\*" $f | tail -3; done
