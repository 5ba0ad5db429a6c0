#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm_tma.py tests/test_gpu_nrms.py tests/test_gpu_fullsize_properties.py -q -x -p no:cacheprovider 2>&1 | tail -3
run() {  # label, env...
  label=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench23_$label.json 2> gpurun_out/bench23_$label.err
  python - $label <<'PY'
import json,sys
d=json.loads(open(f"gpurun_out/bench23_{sys.argv[1]}.json").read())
k=d["kernel_ms_per_step"]
print(sys.argv[1],round(d["value"]), d["ms_per_step_repeats"], "e2e", round(d["e2e"]["value"]), {x:k[x] for x in ("news.qkv_wgrad_gemm","news.adam")}, d["clocks"]["sm_mhz"])
PY
}
run new A=1
run old EBK_GEMM_BN_BIGK=0
run new2 A=1

