#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
EBK_DP_OUT=/tmp timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/dp_fit_check.py > gpurun_out/dp_fit_check.log 2>&1
grep -E "DP_FIT_CHECK|rank [01]:|Error|error" gpurun_out/dp_fit_check.log | head -10
timeout 900 python -m pytest tests/test_gpu_fit_surface.py -q -p no:cacheprovider -k "data_parallel" 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_dp2b.json 2> gpurun_out/bench_dp2b.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/bench_dp2b.json").read().strip().splitlines()[-1])
    print("N=2", round(d["value"]), d["ms_per_step_repeats"], "e2e", round(d["e2e"]["value"]), "comm_exposed", d.get("comm_ms_exposed"), d.get("comm_ms"))
except Exception as e: print("ERR", e)
PY
tail -2 gpurun_out/bench_dp2b.err
