#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,launch__grid_size,launch__registers_per_thread
timeout 900 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/r02_docvec_step_metrics.csv python tools/profile_step.py --workload docvec_bs512 > gpurun_out/ncu_docvec.log 2>&1
tail -1 gpurun_out/ncu_docvec.log; wc -l gpurun_out/r02_docvec_step_metrics.csv
