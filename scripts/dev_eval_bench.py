"""Developer micro-benchmark of ci_kalman_logp_grad (not the contract bench)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, torch
from tfcausalimpact_b200.engine import DeviceBatch

def run(N, G, T, Tf, k, S, iters=10, warm=3):
    rng = np.random.default_rng(0)
    y = rng.normal(size=(N, T)).astype(np.float32)
    X = rng.normal(size=(N, T + Tf, k)).astype(np.float32) if k else None
    db = DeviceBatch(y, X, Tf=Tf, nseasons=S)
    P = db.P
    u = (torch.randn(P, N * G, device='cuda') * 0.3)
    lp = torch.empty(N * G, device='cuda'); gr = torch.empty_like(u)
    for _ in range(warm): db.logp_grad(u, lp, gr)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): db.logp_grad(u, lp, gr)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    d = S
    w = 3 * (5 * d * d + 6 * d + 2 * k + 16) * T * N * G
    print(f'N={N} G={G} T={T} k={k} S={S}: {ms:.3f} ms/eval  {N*G/ms*1e3/1e6:.2f} M pair-evals/s  '
          f'{w/ms/1e9:.2f} algorithmic TFLOP/s  finite={bool(torch.isfinite(lp).all())}')

if __name__ == '__main__':
    it = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    run(1000, 8, 365, 30, 10, 7, it)
    run(10000, 8, 365, 30, 50, 7, it)
    run(10000, 8, 365, 30, 0, 7, it)
    run(10000, 8, 365, 30, 0, 1, it)
    run(1, 1, 90, 18, 3, 1, it)
