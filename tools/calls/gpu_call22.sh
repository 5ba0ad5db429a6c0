#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
run() {  # label, env...
  label=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench22_$label.json 2> gpurun_out/bench22_$label.err
  python - $label <<'PY'
import json,sys
d=json.loads(open(f"gpurun_out/bench22_{sys.argv[1]}.json").read())
k=d["kernel_ms_per_step"]
print(sys.argv[1],round(d["value"]), d["ms_per_step_repeats"], "e2e", round(d["e2e"]["value"]), {x:k[x] for x in ("news.qkv_wgrad_gemm","news.adam")}, d["clocks"]["sm_mhz"])
PY
}
run base A=1
run bn208 EBK_GEMM_BN_BIGK=208
run bn176 EBK_GEMM_BN_BIGK=176
run bn160 EBK_GEMM_BN_BIGK=160
run bn128 EBK_GEMM_BN_BIGK=128
run nodefer EBK_DEFER_WGRAD=0
run bn208nodefer EBK_GEMM_BN_BIGK=208 EBK_DEFER_WGRAD=0
