#!/bin/bash
# round 2, GPU call 4: suite, bench A/B (fused / unfused / nopair), ncu of one step
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/gputest4.log 2>&1
echo "pytest rc=$?" >> gpurun_out/gputest4.log
grep -E "passed|failed|^FAILED|rc=" gpurun_out/gputest4.log | tail -20
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench4_n1.json 2> gpurun_out/bench4_n1.err
EBK_FUSED_ATTN=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench4_n1_unfused.json 2> gpurun_out/bench4_n1_unfused.err
EBK_FUSED_PAIR=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench4_n1_nopair.json 2> gpurun_out/bench4_n1_nopair.err
for f in gpurun_out/bench4_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    print(round(d["value"]), d["ms_per_step_repeats"], "e2e", round(d["e2e"]["value"]), d["e2e"]["ms_per_step_repeats"], d["kernel_ms_per_step"])
except Exception as e: print("ERR", e)
PY
done
tail -3 gpurun_out/bench4_n1.err
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -f -o gpurun_out/r02_step python tools/profile_step.py --no-defer > gpurun_out/ncu_step.log 2>&1
tail -3 gpurun_out/ncu_step.log; ls -la gpurun_out/*.ncu-rep
