#!/bin/bash
# compute-sanitizer over the kernels changed in the second half of round 2 (small cases)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $CS --tool memcheck --error-exitcode 9 python -m pytest "tests/test_gpu_nrms.py::test_loss_and_gradients[case1-0.2-1]" "tests/test_gpu_docvec.py::test_docvec_loss_gradients_and_bn_stats[case1-0.2-1]" tests/test_gpu_docvec.py::test_docvec_graph_replay_matches_eager_steps -q -x -p no:cacheprovider > gpurun_out/sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|invalid" gpurun_out/sanitizer_memcheck.log | tail -5
timeout 1200 $CS --tool racecheck --racecheck-report all --error-exitcode 9 python -m pytest "tests/test_gpu_nrms.py::test_loss_and_gradients[case1-0.2-1]" "tests/test_gpu_docvec.py::test_docvec_loss_gradients_and_bn_stats[case1-0.2-1]" -q -x -p no:cacheprovider > gpurun_out/sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer_racecheck.log | tail -5
grep -E "Race reported|hazard" gpurun_out/sanitizer_racecheck.log | sed 's/0x[0-9a-f]*//g' | sort | uniq -c | sort -rn | head -20
