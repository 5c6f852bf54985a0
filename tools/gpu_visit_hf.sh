#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_hf_llama.py tests/test_gpu_cache.py tests/test_gpu_round2.py -m gpu -q -p no:cacheprovider -x --tb=short > gpurun_out/tests_hf.log 2>&1; tail -25 gpurun_out/tests_hf.log
timeout 600 python tools/speedtest.py --layers 8 --prefill 1024 16384 --decode 48 --niter 1 > gpurun_out/speedtest.log 2>&1; tail -4 gpurun_out/speedtest.log | cut -c1-900
