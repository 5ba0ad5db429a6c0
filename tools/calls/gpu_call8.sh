#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fullsize_properties.py tests/test_gpu_fit_surface.py -q -p no:cacheprovider 2>&1 | tail -12
timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:gemm_tma_kernel -s 8 -c 1 -f -o gpurun_out/r02_att_dgrad python tools/profile_step.py --no-defer > gpurun_out/ncu_attdgrad.log 2>&1
tail -2 gpurun_out/ncu_attdgrad.log
ls -la gpurun_out/*.ncu-rep
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench8_n1.json 2> gpurun_out/bench8_n1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench8_n1.json").read())
print(round(d["value"]), d["ms_per_step_repeats"], "e2e", round(d["e2e"]["value"]), d["e2e"]["ms_per_step_repeats"], d["kernel_ms_per_step"])
PY
