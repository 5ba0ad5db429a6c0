#!/bin/bash
# round 2 final single-GPU records: GPU tests, smoke, bench (default + other workloads + reference arm), ncu metric list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/gputest_final.log 2>&1
echo "pytest rc=$?" >> gpurun_out/gputest_final.log
grep -E "passed|failed|^FAILED|rc=" gpurun_out/gputest_final.log | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_final_reference.json 2> gpurun_out/bench_final_reference.err
for wl in nrms_ebnerd_large_shape_h50_bs256 docvec_bs512 naml_h50_bs64 nrms_ebnerd_small_xlmr_large_bs256 nrms_dummy_bs32; do
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload $wl > gpurun_out/bench_final_$wl.json 2> gpurun_out/bench_final_$wl.err
done
for f in gpurun_out/bench_final_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    e=d.get("e2e",{})
    print(round(d["value"],1), d.get("ms_per_step_repeats"), "e2e", round(e.get("value",0),1), e.get("ms_per_step_repeats"), "scorer", (d.get("scorer") or {}).get("candidates_per_s"), "launches", d.get("gpu_launches"))
except Exception as ex: print("ERR", ex)
PY
done
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,sm__throughput.avg.pct_of_peak_sustained_elapsed
timeout 900 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/r02_step_metrics.csv python tools/profile_step.py --no-defer > gpurun_out/ncu_step.log 2>&1
tail -1 gpurun_out/ncu_step.log; wc -l gpurun_out/r02_step_metrics.csv
timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:gemm_tma_kernel -c 1 -f -o gpurun_out/r02_fused16 python tools/profile_step.py --no-defer > gpurun_out/ncu_fused.log 2>&1
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out
