#!/bin/bash
mkdir -p gpurun_out
MILLION_B200_LIB=$PWD/variants/debug.so python tools/attn_phases.py 8 32768 > gpurun_out/phases_bs8.log 2>&1
MILLION_B200_LIB=$PWD/variants/debug.so python tools/attn_phases.py 1 32768 > gpurun_out/phases_bs1.log 2>&1
MILLION_B200_LIB=$PWD/variants/debug.so python tools/attn_phases.py 1 16384 > gpurun_out/phases_bs1_16k.log 2>&1
python tools/graph_rate.py 32768 1 8 2>&1 | grep "pdl 1" > gpurun_out/rate_now.log
cat gpurun_out/phases_bs8.log gpurun_out/phases_bs1.log gpurun_out/phases_bs1_16k.log gpurun_out/rate_now.log
