#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
timeout 1200 python -m pytest tests/test_gpu_docvec.py tests/test_gpu_nrms_dense.py tests/test_gpu_naml.py tests/test_gpu_reference_golden.py tests/test_gpu_fit_surface.py -q -x -p no:cacheprovider 2>&1 | tail -5
for t in 1 0; do
EBK_DENSE_TMA=$t timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload docvec_bs512 > gpurun_out/bench17_docvec_tma$t.json 2> gpurun_out/bench17_docvec_tma$t.err
python - $t <<'PY'
import json,sys
try:
    d=json.loads(open(f"gpurun_out/bench17_docvec_tma{sys.argv[1]}.json").read())
    print("docvec tma",sys.argv[1],round(d["value"]), d["ms_per_step_repeats"], "e2e", round(d["e2e"]["value"]), d["e2e"].get("ms_per_step_repeats"), "launches", d.get("gpu_launches"))
except Exception as ex:
    print("ERR", ex); print(open(f"gpurun_out/bench17_docvec_tma{sys.argv[1]}.err").read()[-1500:])
PY
done
M=gpu__time_duration.sum,launch__grid_size
timeout 900 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/r02_docvec_step_metrics3.csv python tools/profile_step.py --workload docvec_bs512 > gpurun_out/ncu_docvec.log 2>&1
tail -1 gpurun_out/ncu_docvec.log
