#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
run() {  # n csr
timeout 900 env EBK_DP_CSR_GATHER=$2 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 2953$1 bench.py --gpus $1 --steps 20 --warmup 5 > gpurun_out/bench_dp$1_csr$2.json 2> gpurun_out/bench_dp$1_csr$2.err
python - $1 $2 <<'PY'
import json,sys
n,c=sys.argv[1:3]
try:
    d=json.loads(open(f"gpurun_out/bench_dp{n}_csr{c}.json").read().strip().splitlines()[-1])
    k=d["kernel_ms_per_step"]
    print("N=",n,"csr=",c, round(d["value"]), d["ms_per_step_repeats"], "e2e", round(d["e2e"]["value"]), "gather", k.get("news.embed_gather"), "exposed", d.get("comm_ms_exposed"), d.get("comm_ms"))
except Exception as e: print("ERR", e)
PY
}
run 8 1
run 8 0
run 4 1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tools/dp_check.py 2>&1 | grep DP_CHECK
