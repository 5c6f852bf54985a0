#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -q -k "fused_splitkv or fuzz or fast_kernels" -p no:cacheprovider > gpurun_out/tests_fused.log 2>&1; tail -3 gpurun_out/tests_fused.log
bash tools/gpu_visit_n.sh $1
