#!/bin/bash
# round 2, GPU call 6: 16-warp fused epilogue, fit-based e2e
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_nrms.py -k "fused_projection" -q -p no:cacheprovider 2>&1 | tail -5
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_nrms.py -k "fused_projection and shape0" -q -x -p no:cacheprovider 2>&1 | tail -4
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_nrms.py -k "fused_projection and shape1" -q -x -p no:cacheprovider 2>&1 | tail -6
timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/gputest6.log 2>&1
echo "pytest rc=$?" >> gpurun_out/gputest6.log
grep -E "passed|failed|^FAILED|rc=" gpurun_out/gputest6.log | tail -20
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench6_n1.json 2> gpurun_out/bench6_n1.err
EBK_FUSED_ATTN=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench6_n1_unfused.json 2> gpurun_out/bench6_n1_unfused.err
EBK_FUSED_PAIR=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench6_n1_nopair.json 2> gpurun_out/bench6_n1_nopair.err
for f in gpurun_out/bench6_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    print(round(d["value"]), d["ms_per_step_repeats"], "e2e", round(d["e2e"]["value"]), d["e2e"]["ms_per_step_repeats"], "tob", round(d["e2e"]["train_on_batch_sync"]["value"]), d["kernel_ms_per_step"])
except Exception as e: print("ERR", e)
PY
done
tail -3 gpurun_out/bench6_n1.err
