#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 120 python tools/debug_gemm_case.py 768 1200 3841 0 0 255 2>&1 | tail -12
timeout 120 python tools/debug_gemm_case.py 768 1200 3840 0 0 255 2>&1 | tail -6
timeout 120 python tools/debug_gemm_case.py 768 1200 3841 0 1 255 2>&1 | tail -6
timeout 600 python -m pytest tests/test_gpu_gemm_tma.py -q -p no:cacheprovider 2>&1 | tail -12
M=gpu__time_duration.sum,launch__grid_size
timeout 900 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/r02_naml_step_metrics.csv python tools/profile_step.py --workload naml_h50_bs64 > gpurun_out/ncu_naml.log 2>&1
tail -1 gpurun_out/ncu_naml.log
