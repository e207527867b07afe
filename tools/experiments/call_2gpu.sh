#!/bin/bash
# `chiron call` on the bundled reads: 1 GPU vs 2 ranks under torchrun (reads sharded, no collective): identical result files?
set -e
IN=tests/golden/DNA/raw
rm -rf /tmp/o1 /tmp/o2
python -m chiron_b200.chiron_eval -i $IN -o /tmp/o1 -m DNA_default -p dna-pre --beam 0 --precision tc > /tmp/o1.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 -m chiron_b200.chiron_eval -i $IN -o /tmp/o2 -m DNA_default -p dna-pre --beam 0 --precision tc > /tmp/o2.log 2>&1
ls /tmp/o2/result /tmp/o2/meta
for f in /tmp/o1/result/*; do cmp $f /tmp/o2/result/$(basename $f) && echo "identical: $(basename $f)"; done
tail -2 /tmp/o1.log; tail -2 /tmp/o2.log
