#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
for n in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/bench_dp$n.json 2> gpurun_out/bench_dp$n.err
python - $n <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/bench_dp{n}.json").read().strip().splitlines()[-1])
    print("N=",n, round(d["value"]), d["ms_per_step_repeats"], "e2e", round(d["e2e"]["value"]), "comm_exposed", d.get("comm_ms_exposed"), d.get("comm_ms"))
    print(d["kernel_ms_per_step"])
except Exception as e: print("ERR", e)
PY
tail -2 gpurun_out/bench_dp$n.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tools/dp_check.py 2>&1 | grep DP_CHECK
