"""GPU debugging aid for the tcgen05 GEMM: structured operands that reveal layout mistakes."""
import sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "ebnerd-benchmark_b200")]
from ebrec.models.newsrec import _ebk

lib = _ebk.lib()

def run(tA, tB, M, N, K, A, B, math=1, beta=0.0, C0=None):
    Ad, Bd = torch.tensor(A, dtype=torch.float32).cuda(), torch.tensor(B, dtype=torch.float32).cuda()
    Cd = torch.zeros(M, N, device="cuda") if C0 is None else torch.tensor(C0, dtype=torch.float32).cuda()
    _ebk.check(lib.ebk_gemm(math, tA, tB, M, N, K, _ebk.ptr(Ad), A.shape[1], _ebk.ptr(Bd), B.shape[1], _ebk.ptr(Cd), N, beta, _ebk.stream()))
    torch.cuda.synchronize()
    return Cd.cpu().numpy()

def probe(tA, tB, M, N, K):
    # A(m,k) = (m % 16) * 64 + k  (exact in tf32 for K<=64); B(k,n) = 1 if k == n % K
    Am = np.array([[(m % 16) * 64 + k for k in range(K)] for m in range(M)], dtype=np.float32)
    Bm = np.array([[1.0 if k == (n % K) else 0.0 for n in range(N)] for k in range(K)], dtype=np.float32)
    want = Am @ Bm
    A = Am.T.copy() if tA else Am
    B = Bm.T.copy() if tB else Bm
    got = run(tA, tB, M, N, K, A, B)
    ok = (got == want)
    print(f"tA={tA} tB={tB} M={M} N={N} K={K}: exact match {ok.mean():.3f}")
    if not ok.all():
        bad = np.argwhere(~ok)[:12]
        for m, n in bad:
            g = got[m, n]
            print(f"   C[{m},{n}] got {g} (m%16={int(g)//64}, k={int(g)%64}) want {want[m,n]} (m%16={m%16}, k={n%K})")
        print("   row0 got :", got[0, :40].astype(int).tolist())
        print("   row0 want:", want[0, :40].astype(int).tolist())
        print("   col0 got :", got[:40, 0].astype(int).tolist())
        print("   col0 want:", want[:40, 0].astype(int).tolist())
    return ok.all()

if __name__ == "__main__":
    torch.cuda.set_device(0)
    for (tA, tB) in [(0, 1), (0, 0), (1, 1), (1, 0)]:
        for (M, N, K) in [(128, 64, 32), (128, 256, 32), (128, 256, 64), (256, 512, 128)]:
            try:
                probe(tA, tB, M, N, K)
            except Exception as e:
                print("ERR", tA, tB, M, N, K, e)
    rng = np.random.default_rng(0)
    for (tA, tB) in [(0, 1), (0, 0), (1, 1), (1, 0)]:
        M, N, K = 300, 1200, 768
        A = rng.standard_normal((K, M) if tA else (M, K)).astype(np.float32)
        B = rng.standard_normal((N, K) if tB else (K, N)).astype(np.float32)
        want = (A.T if tA else A).astype(np.float64) @ (B.T if tB else B).astype(np.float64)
        got = run(tA, tB, M, N, K, A, B)
        print(f"random tA={tA} tB={tB}: max err/sqrt(K) = {np.abs(got-want).max()/np.sqrt(K):.3e}")
