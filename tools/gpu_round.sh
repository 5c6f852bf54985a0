#!/bin/bash
# One GPU-box visit: GPU tests, smoke, the N=1 bench (both arms).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1800 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -s > gpurun_out/tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?" >> gpurun_out/bench_n1.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
grep -a "FAILED\|passed\|failed" gpurun_out/tests.log | tail -15; tail -2 gpurun_out/smoke.log; tail -c 300 gpurun_out/bench_n1.err
