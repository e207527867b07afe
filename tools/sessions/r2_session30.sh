#!/bin/bash
# same-box A/B: beam search without (libA = HEAD) / with (working tree) the quiet-frame test
out=gpurun_out/r02_s30; mkdir -p $out
for rep in 1 2; do
echo "-- A (no quiet-frame test)" >> $out/ab.txt
CHIRON_B200_LIB=ab_libs/libA.so timeout 300 python tools/experiments/beam_real_ab.py 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['B'],d['L'],d['beam'],'default',d['modes']['default']['ms'],'pool_24W',d['modes']['pool_24W']['ms'])" >> $out/ab.txt
echo "-- working tree (quiet-frame test)" >> $out/ab.txt
timeout 300 python tools/experiments/beam_real_ab.py 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['B'],d['L'],d['beam'],'default',d['modes']['default']['ms'],'pool_24W',d['modes']['pool_24W']['ms'])" >> $out/ab.txt
done
cat $out/ab.txt
