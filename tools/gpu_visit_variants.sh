#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/variants.log
for rep in 1 2; do
  python tools/graph_rate.py 32768 1 8 2>&1 | grep "pdl 1" >> gpurun_out/variants.log
  for v in variants/*.so; do MILLION_B200_LIB=$PWD/$v python tools/graph_rate.py 32768 1 8 2>&1 | grep "pdl 1" >> gpurun_out/variants.log; done
done
cat gpurun_out/variants.log
