#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_nrms.py -q -x -p no:cacheprovider 2>&1 | tail -3
run() {  # label, env...
  label=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench12_$label.json 2> gpurun_out/bench12_$label.err
  python - $label <<'PY'
import json,sys
d=json.loads(open(f"gpurun_out/bench12_{sys.argv[1]}.json").read())
k=d["kernel_ms_per_step"]
print(sys.argv[1],round(d["value"]), d["ms_per_step_repeats"], "e2e", round(d["e2e"]["value"]), {x:k[x] for x in ("news.embed_gather","news.attn_core_bwd","news.att_dgrad_gemm","news.qkv_wgrad_gemm","news.qkv_dgrad_gemm")}, d["clocks"]["sm_mhz"])
PY
}
run base A=1
run oldgather EBK_EMBED_WARP_ROWS=0
run stages2 EBK_ATT_STAGES=2
run minb4 EBK_ATT_MINB=4
run base2 A=1
cap() {  # name regex skip
  timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:$2 -s $3 -c 1 -f -o gpurun_out/r02_$1 python tools/profile_step.py --no-defer > gpurun_out/ncu_$1.log 2>&1
  tail -1 gpurun_out/ncu_$1.log
}
cap att_dgrad3 gemm_tma_kernel 8
cap embed_warp embed_rows_warp_kernel 0
