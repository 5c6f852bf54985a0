#!/bin/bash
mkdir -p gpurun_out
python tools/precision_budget.py > gpurun_out/precision_default.log 2>&1
MILLION_B200_LIB=$PWD/variants/flush1.so python tools/precision_budget.py > gpurun_out/precision_flush1.log 2>&1
for rep in 1 2; do
  python tools/graph_rate.py 32768 1 8 >> gpurun_out/graph_rate.log 2>&1
  MILLION_B200_LIB=$PWD/variants/pfmerge.so python tools/graph_rate.py 32768 1 8 >> gpurun_out/graph_rate.log 2>&1
  MILLION_B200_LIB=$PWD/variants/flush1.so python tools/graph_rate.py 32768 8 >> gpurun_out/graph_rate.log 2>&1
done
timeout 600 python tools/speedtest.py --layers 4 --prefill 1024 8192 --decode 32 --niter 1 > gpurun_out/speedtest.log 2>&1
cat gpurun_out/precision_default.log gpurun_out/precision_flush1.log gpurun_out/graph_rate.log; tail -5 gpurun_out/speedtest.log
