#!/bin/bash
# Build the library of a git revision (default HEAD) into build/ab/lib<tag>.so for same-box A/B timing:
#   tools/ab_build.sh A [rev];  then on the GPU box:  CHIRON_B200_LIB=build/ab/libA.so python tools/gpu_quick.py ...
set -e
tag=${1:-A}; rev=${2:-HEAD}
root=$(cd "$(dirname "$0")/.." && pwd)
wt=/tmp/ab_$tag
rm -rf $wt; git -C $root worktree prune; git -C $root worktree add -f --detach $wt $rev > /dev/null
make -C $wt/chiron_b200/csrc -j8 > /dev/null
mkdir -p $root/build/ab; cp $wt/chiron_b200/lib/libchiron_b200.so $root/build/ab/lib$tag.so
git -C $root worktree remove --force $wt
echo built build/ab/lib$tag.so from $rev
