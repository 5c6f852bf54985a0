#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/variants.log
for rep in 1 2; do
  for bs in 8 4 2 1; do python tools/graph_rate.py 32768 $bs 2>&1 | grep "pdl 1" >> gpurun_out/variants.log; done
  for v in variants/direct8.so variants/direct12.so variants/direct20.so; do for bs in 8 4 2 1; do MILLION_B200_LIB=$PWD/$v python tools/graph_rate.py 32768 $bs 2>&1 | grep "pdl 1" >> gpurun_out/variants.log; done; done
done
sort gpurun_out/variants.log | uniq
