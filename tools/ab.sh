#!/bin/bash
# A/B of library builds on the same box: tools/ab.sh <bs> variants/a.so variants/b.so ...
bs=$1; shift
for rep in 1 2; do for so in "$@"; do echo "== $so"; MILLION_B200_LIB=$PWD/$so timeout 200 python tools/loop_rate.py $bs 0; done; done
