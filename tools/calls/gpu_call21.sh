#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
run() {  # label, env...
  label=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench21_$label.json 2> gpurun_out/bench21_$label.err
  python - $label <<'PY'
import json,sys
d=json.loads(open(f"gpurun_out/bench21_{sys.argv[1]}.json").read())
k=d["kernel_ms_per_step"]
print(sys.argv[1],round(d["value"]), d["ms_per_step_repeats"], {x:k[x] for x in ("news.att_gemm_fwd","news.qkv_wgrad_gemm","news.qkv_dgrad_gemm","news.att_wgrad_gemm")}, d["clocks"]["sm_mhz"])
PY
}
run base A=1
run attfwd_mt1 EBK_ATT_FWD_TALL=0
run wgrad_bn200 EBK_GEMM_BN_BIGK=200
run wgrad_bn208 EBK_GEMM_BN_BIGK=208
run base2 A=1
