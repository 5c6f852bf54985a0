#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_outliers.py tests/test_gpu_round2.py -m gpu -q -p no:cacheprovider -k "outlier" > gpurun_out/tests_out.log 2>&1; tail -3 gpurun_out/tests_out.log
for k in 0 1 2 4; do python tools/prof_attn.py --bs 8 --kout $k --layers 8 --iters 5 2>&1 | tail -1; done
for k in 0 1 2; do python tools/prof_attn.py --bs 16 --ctx 65536 --nh 4 --nhk 4 --kout $k --layers 8 --iters 3 2>&1 | tail -1; done
python tools/enc_rate.py 64
