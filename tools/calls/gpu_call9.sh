#!/bin/bash
# quick check of a kernel change: kernel-level parity tests, the NRMS parity suite, bench with the per-kernel profile
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_gemm_tma.py tests/test_gpu_nrms.py tests/test_gpu_reference_golden.py -q -x -p no:cacheprovider 2>&1 | tail -6
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench9_n1.json 2> gpurun_out/bench9_n1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench9_n1.json").read())
print(round(d["value"]), d["ms_per_step_repeats"], "e2e", round(d["e2e"]["value"]), d["e2e"]["ms_per_step_repeats"])
print(d["kernel_ms_per_step"])
print(d["clocks"])
PY
