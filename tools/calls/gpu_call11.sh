#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_gemm_tma.py tests/test_gpu_nrms.py tests/test_gpu_reference_golden.py tests/test_gpu_naml.py tests/test_gpu_docvec.py -q -x -p no:cacheprovider 2>&1 | tail -6
for ew in 16 8; do
EBK_GEMM_EPI_WARPS=$ew timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench11_ew$ew.json 2> gpurun_out/bench11_ew$ew.err
python - $ew <<'PY'
import json,sys
d=json.loads(open(f"gpurun_out/bench11_ew{sys.argv[1]}.json").read())
k=d["kernel_ms_per_step"]
print("ew",sys.argv[1],round(d["value"]), d["ms_per_step_repeats"], "e2e", round(d["e2e"]["value"]), {x:k[x] for x in ("news.att_dgrad_gemm","news.attpool_fwd","news.attpool_bwd","news.qkv_gemm_fwd","news.att_gemm_fwd","user.att_dgrad_gemm")}, d["clocks"]["sm_mhz"])
PY
done
