#!/bin/bash
# Same-box A/B: alternate the baseline build(s) and the working-tree build.  Usage: tools/ab_run.sh [tags...]
for rep in 1 2; do
  for tag in "$@"; do echo "== lib$tag"; CB_TC_PREFETCH=0 CHIRON_B200_LIB=build/ab/lib$tag.so timeout 100 python tools/gpu_quick.py tc 4096 512 2>&1 | tail -1; done
  echo "== working tree"; CB_TC_PREFETCH=0 timeout 100 python tools/gpu_quick.py tc 4096 512 2>&1 | tail -1
done
