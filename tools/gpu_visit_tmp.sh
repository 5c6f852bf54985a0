#!/bin/bash
mkdir -p gpurun_out /tmp/ncu
for bs in 8 1; do
  ncu --section SourceCounters --clock-control none -k regex:attn_fast_kernel -s 40 -c 1 -f -o /tmp/ncu/h$bs python tools/graph_rate.py 32768 $bs > gpurun_out/ncu_h$bs.log 2>&1
  ncu -i /tmp/ncu/h$bs.ncu-rep --page source --csv --print-source sass > /tmp/ncu/h$bs.csv 2>/dev/null
  python tools/ncu_src_top.py /tmp/ncu/h$bs.csv 70 > gpurun_out/r02_src_fast_bs$bs.txt 2>&1
  python tools/ncu_src_dump.py /tmp/ncu/h$bs.csv > gpurun_out/r02_src_fast_bs${bs}_all.txt 2>&1
done
