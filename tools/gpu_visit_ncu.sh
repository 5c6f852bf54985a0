#!/bin/bash
# ncu evidence of the round: launch list of the bench command, full captures of the headline kernel (batch 8 and 1) and of the
# M=32 kernel.  Reports are summarised on the box (tools/ncu_summary.py) and deleted: gpurun_out/ may not exceed 64 MiB.
mkdir -p gpurun_out /tmp/ncu
CMD="python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline"
$CMD > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv $CMD > gpurun_out/ncu_bench.log 2>&1
cap() {   # name, kernel regex, command...
  local name=$1 rx=$2; shift 2
  "$@" > gpurun_out/plain_$name.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$rx -s 3 -c 1 -f -o /tmp/ncu/$name "$@" > gpurun_out/ncu_$name.log 2>&1
  python tools/ncu_summary.py /tmp/ncu/$name.ncu-rep 25 > gpurun_out/r02_prof_$name.txt 2>&1
  ncu -i /tmp/ncu/$name.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv, sys
rows = list(csv.reader(sys.stdin)); hdr = rows[0]
for r in rows[2:]:
    d = dict(zip(hdr, r)); print({k: d[k] for k in ('Kernel Name', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__time_duration.sum') if k in d})
" > gpurun_out/r02_traffic_$name.txt 2>&1
}
cap fast_bs8 attn_fast_kernel python tools/prof_attn.py --bs 8 --iters 2 --layers 2
cap fast_bs1 attn_fast_kernel python tools/prof_attn.py --bs 1 --iters 2 --layers 2
cap dm4_mha attn_fast_dm4_kernel python tools/prof_attn.py --bs 16 --ctx 65536 --nh 4 --nhk 4 --M 32 --kout 2 --iters 2 --layers 2
ls -la /tmp/ncu; head -30 gpurun_out/r02_prof_fast_bs8.txt; cat gpurun_out/r02_traffic_*.txt
