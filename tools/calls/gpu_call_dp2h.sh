#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_nrms.py -q -x -p no:cacheprovider -k "peer_table or fused_embedding or graph" 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tools/dp_check.py 2>&1 | grep -E "DP_CHECK|Error|error" | head -5
EBK_DP_OUT=/tmp timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 tools/dp_fit_check.py > gpurun_out/dp_fit_check.log 2>&1
grep -E "DP_FIT_CHECK|Error|error" gpurun_out/dp_fit_check.log | head -5 | cut -c1-400
for mode in 1 0; do
EBK_DP_CSR_GATHER=$mode timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2956$mode bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_dp2_csr$mode.json 2> gpurun_out/bench_dp2_csr$mode.err
python - $mode <<'PY'
import json,sys
m=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/bench_dp2_csr{m}.json").read().strip().splitlines()[-1])
    k=d["kernel_ms_per_step"]
    print("N=2 csr=",m, round(d["value"]), d["ms_per_step_repeats"], "e2e", round(d["e2e"]["value"]), "gather", k.get("news.embed_gather"), "scatter", k.get("news.embed_scatter"), "exposed", d.get("comm_ms_exposed"))
except Exception as e: print("ERR", e)
PY
tail -2 gpurun_out/bench_dp2_csr$mode.err | cut -c1-300
done
