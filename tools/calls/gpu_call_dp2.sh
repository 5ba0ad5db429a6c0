#!/bin/bash
# round 2, 2-GPU call: data-parallel correctness (engine + fit) and the N=2 bench line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_fit_surface.py -q -s -p no:cacheprovider -k "data_parallel" 2>&1 | tail -15
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_dp2.json 2> gpurun_out/bench_dp2.err
tail -3 gpurun_out/bench_dp2.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/bench_dp2.json").read().strip().splitlines()[-1])
    print("N=2", round(d["value"]), d["ms_per_step_repeats"], "e2e", round(d["e2e"]["value"]), "comm_exposed", d.get("comm_ms_exposed"), d.get("comm_ms"))
    print(d["kernel_ms_per_step"]); print(d["kernel_roofline_frac"])
except Exception as e: print("ERR", e)
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench_dp2_ref.json 2> gpurun_out/bench_dp2_ref.err
head -c 600 gpurun_out/bench_dp2_ref.json
