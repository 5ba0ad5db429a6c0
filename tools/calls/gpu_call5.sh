#!/bin/bash
# round 2, GPU call 5: 8-warp fused epilogue: focused test, suite, bench A/B, compact ncu metric list of one step
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_nrms.py -k "fused_projection" -q -p no:cacheprovider 2>&1 | tail -5
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_nrms.py -k "fused_projection and shape0" -q -x -p no:cacheprovider 2>&1 | tail -4
timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/gputest5.log 2>&1
echo "pytest rc=$?" >> gpurun_out/gputest5.log
grep -E "passed|failed|^FAILED|rc=" gpurun_out/gputest5.log | tail -20
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench5_n1.json 2> gpurun_out/bench5_n1.err
EBK_FUSED_ATTN=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench5_n1_unfused.json 2> gpurun_out/bench5_n1_unfused.err
EBK_FUSED_PAIR=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench5_n1_nopair.json 2> gpurun_out/bench5_n1_nopair.err
EBK_ATTPOOL_FAST=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench5_n1_oldpool.json 2> gpurun_out/bench5_n1_oldpool.err
for f in gpurun_out/bench5_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    print(round(d["value"]), d["ms_per_step_repeats"], "e2e", round(d["e2e"]["value"]), d["e2e"]["ms_per_step_repeats"], d["kernel_ms_per_step"])
except Exception as e: print("ERR", e)
PY
done
tail -3 gpurun_out/bench5_n1.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,sm__throughput.avg.pct_of_peak_sustained_elapsed
timeout 900 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/r02_step_metrics.csv python tools/profile_step.py --no-defer > gpurun_out/ncu_step.log 2>&1
tail -2 gpurun_out/ncu_step.log; wc -l gpurun_out/r02_step_metrics.csv
timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:gemm_tma_kernel -c 2 -f -o gpurun_out/r02_fused python tools/profile_step.py --no-defer > gpurun_out/ncu_fused.log 2>&1
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out
