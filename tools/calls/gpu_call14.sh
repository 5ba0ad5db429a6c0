#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
timeout 900 python -m pytest tests/test_gpu_docvec.py tests/test_gpu_fit_surface.py tests/test_gpu_nrms_dense.py -q -x -p no:cacheprovider 2>&1 | tail -5
timeout 600 python -m pytest tests/test_gpu_nrms.py -q -x -p no:cacheprovider -k "graph or fused_embedding" 2>&1 | tail -3
for g in 0 1; do
EBK_NO_GRAPH=$g timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload docvec_bs512 > gpurun_out/bench14_docvec_nograph$g.json 2> gpurun_out/bench14_docvec_nograph$g.err
python - $g <<'PY'
import json,sys
try:
    d=json.loads(open(f"gpurun_out/bench14_docvec_nograph{sys.argv[1]}.json").read())
    print("docvec EBK_NO_GRAPH=",sys.argv[1],round(d["value"]), d["ms_per_step_repeats"], "e2e", round(d["e2e"]["value"]), d["e2e"].get("ms_per_step_repeats"), "launches", d.get("gpu_launches"))
except Exception as ex:
    print("ERR", ex); print(open(f"gpurun_out/bench14_docvec_nograph{sys.argv[1]}.err").read()[-1500:])
PY
done
