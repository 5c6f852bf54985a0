#!/bin/bash
# A/B of two source TREES on the same box (different ABI versions cannot share _lib.py): tools/ab_tree.sh <bs> <tree> [<tree> ...]
# a tree is a directory holding a built million_b200/ package ("." = this repo, variants/r1 = the round-1 build)
bs=$1; shift
for rep in 1 2; do for t in "$@"; do echo "== $t"; MILLION_PKG_ROOT=$(realpath $t) timeout 300 python tools/loop_rate.py $bs 0; done; done
