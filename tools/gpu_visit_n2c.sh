#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_round2.py -m gpu -q -k "every_device" -p no:cacheprovider > gpurun_out/tests_2dev.log 2>&1; tail -3 gpurun_out/tests_2dev.log
bash tools/gpu_visit_n.sh 2
