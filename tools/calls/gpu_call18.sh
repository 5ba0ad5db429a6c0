#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
timeout 1500 python -m pytest tests/test_gpu_gemm_tma.py tests/test_gpu_docvec.py tests/test_gpu_nrms_dense.py tests/test_gpu_naml.py tests/test_gpu_reference_golden.py tests/test_gpu_nrms.py -q -x -p no:cacheprovider 2>&1 | tail -5
for t in 1 0; do
EBK_GEMM_SMALL_TILES=$t timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload docvec_bs512 > gpurun_out/bench18_docvec_small$t.json 2> gpurun_out/bench18_docvec_small$t.err
EBK_GEMM_SMALL_TILES=$t timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench18_nrms_small$t.json 2> gpurun_out/bench18_nrms_small$t.err
EBK_GEMM_SMALL_TILES=$t timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload nrms_dummy_bs32 > gpurun_out/bench18_dummy_small$t.json 2> gpurun_out/bench18_dummy_small$t.err
python - $t <<'PY'
import json,sys
for w in ("docvec","nrms","dummy"):
    try:
        d=json.loads(open(f"gpurun_out/bench18_{w}_small{sys.argv[1]}.json").read())
        k=d.get("kernel_ms_per_step") or {}
        print(w,"small",sys.argv[1],round(d["value"]), d["ms_per_step_repeats"], "e2e", round(d["e2e"]["value"]), "user kernels", round(sum(v for n,v in k.items() if n.startswith("user.")),4))
    except Exception as ex:
        print("ERR", w, ex)
PY
done
