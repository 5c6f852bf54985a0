#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/variants_dm4.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "encode" -p no:cacheprovider > gpurun_out/tests_enc.log 2>&1; tail -2 gpurun_out/tests_enc.log
for rep in 1 2; do
  RATE_M=32 python tools/graph_rate.py 32768 8 2>&1 | grep "pdl 1" >> gpurun_out/variants_dm4.log
  RATE_M=32 RATE_NH=4 RATE_NHK=4 python tools/graph_rate.py 65536 16 2>&1 | grep "pdl 1" >> gpurun_out/variants_dm4.log
  for v in variants/dm4*.so; do
    RATE_M=32 MILLION_B200_LIB=$PWD/$v python tools/graph_rate.py 32768 8 2>&1 | grep "pdl 1" >> gpurun_out/variants_dm4.log
    RATE_M=32 RATE_NH=4 RATE_NHK=4 MILLION_B200_LIB=$PWD/$v python tools/graph_rate.py 65536 16 2>&1 | grep "pdl 1" >> gpurun_out/variants_dm4.log
  done
done
sort gpurun_out/variants_dm4.log | uniq
