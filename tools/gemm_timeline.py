"""Per-step clock64 timeline of CTA 0 of one tcgen05 GEMM inside a real NRMS train step (debugging aid).
usage: gemm_timeline.py <target launch index>   (0 = news qkv fwd, ...)"""
import sys, ctypes as C
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "ebnerd-benchmark_b200")]
from ebrec.models.newsrec import _ebk
from ebrec.models.newsrec._engine import NRMSEngine
import bench
lib = _ebk.lib()
torch.cuda.set_device(0)
w = bench.WORKLOADS[bench.DEFAULT_WORKLOAD]
eng = NRMSEngine(V=w["V"], E=w["E"], T=w["T"], H=w["H"], nh=w["nh"], dh=w["dh"], att=w["att"], dropout=0.2, lr=1e-4, seed=1)
rng = np.random.default_rng(0)
eng.params.p("table").normal_(0, 0.02)
for n in ("news_Wqkv", "news_attW", "user_Wqkv", "user_attW", "news_attq", "user_attq"):
    eng.params.p(n).normal_(0, 0.05)
his, pred, y = bench.synth_batch(rng, w, w["B"])
tok, lab = eng.to_device_batch(his, pred, y)
for _ in range(2):
    eng.train_step_dev(tok, lab, w["B"], w["C"])
torch.cuda.synchronize()
lib.ebk_debug_gemm_timeline.argtypes = [C.c_void_p, C.c_int]
for target in [int(a) for a in sys.argv[1:]] or [0]:
    dbg = torch.zeros(3 * 96 * 4, dtype=torch.int64, device="cuda")
    lib.ebk_debug_gemm_timeline(C.c_void_p(dbg.data_ptr()), target)
    eng.train_step_dev(tok, lab, w["B"], w["C"])
    torch.cuda.synchronize()
    lib.ebk_debug_gemm_timeline(None, 0)
    d = dbg.cpu().numpy().reshape(3, 96, 4)
    t0 = d[0, 0, 0]
    print(f"=== GEMM launch #{target}: producer thread0 per step [loads_issued, empty_ok, sts_done, arrived] | mma [wait, full_ok, issued]")
    for g in list(range(0, 12)) + list(range(40, 60)):
        print(g, *(int(x - t0) for x in d[0, g]), "|", *(int(x - t0) for x in d[1, g][:3]))
    print("epilogue items [wait_start, tfull_ok, done]:", [[int(x - t0) for x in d[2, t][:3]] for t in range(3)])
