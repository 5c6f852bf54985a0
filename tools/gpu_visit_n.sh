#!/bin/bash
# N-GPU bench line (KV-head-sharded headline + split-KV 128K + 7B KV-head extras): tools/gpu_visit_n.sh N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_n$N.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "exit $?" >> gpurun_out/bench_n$N.err
tail -c 1500 gpurun_out/bench_n$N.err; python - <<PY
import json
try:
    j = json.load(open("gpurun_out/bench_n$N.json"))
    print("value", j["value"], "frac", j["roofline"]["frac"], "e2e", j["e2e"]["value"])
    print(json.dumps(j["extra"].get("splitkv_128k"), indent=1))
    print(json.dumps(j["extra"].get("kvhead_7b_64k_bs16"))[:1500])
except Exception as e:
    print("no json:", e)
PY
