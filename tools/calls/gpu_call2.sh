#!/bin/bash
# round 2, GPU call 2: GPU tests after the graph / opts refactor, bench N=1 (graph), other workloads
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/gputest2.log 2>&1
echo "pytest rc=$?" >> gpurun_out/gputest2.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke2.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench2_n1.json 2> gpurun_out/bench2_n1.err
EBK_NO_GRAPH=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench2_n1_nograph.json 2> gpurun_out/bench2_n1_nograph.err
for wl in nrms_ebnerd_large_shape_h50_bs256 docvec_bs512 naml_h50_bs64 nrms_ebnerd_small_xlmr_large_bs256; do
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload $wl > gpurun_out/bench2_$wl.json 2> gpurun_out/bench2_$wl.err
done
grep -E "passed|failed|^FAILED|rc=" gpurun_out/gputest2.log | tail -20; tail -2 gpurun_out/smoke2.log
for f in gpurun_out/bench2_*.json; do echo $f; head -c 400 $f; echo; done
tail -5 gpurun_out/bench2_n1.err
