#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/gputest7.log 2>&1
echo "pytest rc=$?" >> gpurun_out/gputest7.log
grep -E "passed|failed|^FAILED|rc=" gpurun_out/gputest7.log | tail -8
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench7_n1.json 2> gpurun_out/bench7_n1.err
EBK_GEMM_EPI_WARPS=4 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench7_n1_ew4.json 2> gpurun_out/bench7_n1_ew4.err
for f in gpurun_out/bench7_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    print(round(d["value"]), d["ms_per_step_repeats"], "e2e", round(d["e2e"]["value"]), d["e2e"]["ms_per_step_repeats"], "tob", round(d["e2e"]["train_on_batch_sync"]["value"]), d["kernel_ms_per_step"])
except Exception as e: print("ERR", e)
PY
done
tail -3 gpurun_out/bench7_n1.err
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest tests/test_gpu_nrms.py -k "fused_projection and shape1" -q -x -p no:cacheprovider > gpurun_out/racecheck.log 2>&1
grep -E "Race reported|ERROR|WARN|hazard" gpurun_out/racecheck.log | sort | uniq -c | sort -rn | head -20
