#!/bin/bash
# ncu --set full captures of the HBM-side kernels that sit below their roof (one launch each, the news-encoder instance)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
cap() {  # name regex skip
  timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:$2 -s $3 -c 1 -f -o gpurun_out/r02_$1 python tools/profile_step.py --no-defer > gpurun_out/ncu_$1.log 2>&1
  tail -1 gpurun_out/ncu_$1.log
}
cap attpool_bwd attpool_bwd_fused_kernel 1
cap attpool_fwd attpool_fwd_fast_kernel 0
cap attn_bwd attn_bwd_pre_kernel 1
cap embed_rows embed_rows_kernel 0
cap att_dgrad2 gemm_tma_kernel 8
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out
