#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_pq.py tests/test_gpu_round2.py -m gpu -q -p no:cacheprovider -x > gpurun_out/tests_small.log 2>&1
tail -5 gpurun_out/tests_small.log
