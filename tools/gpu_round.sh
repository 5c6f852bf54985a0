#!/bin/bash
# One GPU-box visit: GPU tests, smoke, the N=1 bench (both arms), A/B of the main loop against the round-1 tree.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python tools/dbg_fused.py > gpurun_out/dbg_fused.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -s > gpurun_out/tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?" >> gpurun_out/bench_n1.err
timeout 300 python bench.py --no-extras --no-cpu-baseline --pdl 0 > gpurun_out/bench_n1_nopdl.json 2> gpurun_out/bench_n1_nopdl.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 300 bash tools/ab_tree.sh 8 variants/r1 . > gpurun_out/ab8.log 2>&1
timeout 300 bash tools/ab_tree.sh 1 variants/r1 . > gpurun_out/ab1.log 2>&1
cat gpurun_out/dbg_fused.log; grep -a "FAILED\|passed\|failed" gpurun_out/tests.log | tail -15; tail -2 gpurun_out/smoke.log; tail -c 400 gpurun_out/bench_n1.err; cat gpurun_out/ab8.log gpurun_out/ab1.log
