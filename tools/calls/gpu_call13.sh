#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_nrms.py -q -x -p no:cacheprovider 2>&1 | tail -3
run() {  # label, env...
  label=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench13_$label.json 2> gpurun_out/bench13_$label.err
  python - $label <<'PY'
import json,sys
d=json.loads(open(f"gpurun_out/bench13_{sys.argv[1]}.json").read())
k=d["kernel_ms_per_step"]
print(sys.argv[1],round(d["value"]), d["ms_per_step_repeats"], "e2e", round(d["e2e"]["value"]), {x:k[x] for x in ("news.attn_core_bwd","user.attn_core_bwd","news.qkv_wgrad_gemm","news.qkv_dgrad_gemm")}, d["clocks"]["sm_mhz"])
PY
}
run pf1 A=1
run pf0 EBK_ATT_L2_PREFETCH=0
run pf1b A=1
run pf0b EBK_ATT_L2_PREFETCH=0
