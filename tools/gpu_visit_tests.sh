#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -s > gpurun_out/tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/tests.log
for rep in 1 2; do
  python tools/graph_rate.py 32768 1 8 >> gpurun_out/graph_rate2.log 2>&1
  MILLION_B200_LIB=$PWD/variants/flush2.so python tools/graph_rate.py 32768 1 8 >> gpurun_out/graph_rate2.log 2>&1
done
python tools/prof_attn.py --bs 16 --ctx 65536 --nh 4 --nhk 4 --kout 2 --layers 8 > gpurun_out/prof_7b.log 2>&1
grep -a "FAILED\|passed\|failed" gpurun_out/tests.log | tail -15; cat gpurun_out/graph_rate2.log gpurun_out/prof_7b.log
