#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tools/dp_check.py 2>&1 | grep -E "DP_CHECK|Error|error" | head -5
EBK_DP_OUT=/tmp timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 tools/dp_fit_check.py > gpurun_out/dp_fit_check.log 2>&1
grep -E "DP_FIT_CHECK|Error|error" gpurun_out/dp_fit_check.log | head -5 | cut -c1-600
for mode in 1; do
EBK_DP_PULL=$mode timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2956$mode bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_dp2_pull$mode.json 2> gpurun_out/bench_dp2_pull$mode.err
python - $mode <<'PY'
import json,sys
m=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/bench_dp2_pull{m}.json").read().strip().splitlines()[-1])
    print("N=2 pull=",m, round(d["value"]), d["ms_per_step_repeats"], "e2e", round(d["e2e"]["value"]), "comm_exposed", d.get("comm_ms_exposed"), d.get("comm_ms"), "adam", d["kernel_ms_per_step"].get("news.adam"))
except Exception as e: print("ERR", e)
PY
tail -2 gpurun_out/bench_dp2_pull$mode.err
done
