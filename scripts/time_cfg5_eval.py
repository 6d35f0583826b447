"""Developer timing of the large-latent evaluation at configs[4] shape (100 lanes x 10,000 steps, nseasons 52):
register-resident 4-warp kernel (ci_eval_rw.cu) against the shared-memory CTA-per-lane kernel (k_eval_cta).
usage: python scripts/time_cfg5_eval.py"""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tfcausalimpact_b200 import _native as nt
import torch, bench
from tfcausalimpact_b200.engine import DeviceBatch
N, T, Tf, S = 100, 10000, 520, 52
y, _ = bench.synth(N, T, Tf, 0, S, 5, 'cuda')
db = DeviceBatch(y, None, Tf=Tf, nseasons=S)
u = torch.randn(db.P, N, device='cuda') * 0.3
out = {}
ref = None
for rw, nw in ((0, 4), (1, 4), (1, 8)):
    nt.check(nt.lib().ci_set_option(b'eval_rw', ctypes.c_long(rw)), 'opt')
    nt.check(nt.lib().ci_set_option(b'eval_rw_warps', ctypes.c_long(nw)), 'opt')
    lp, g = db.logp_grad(u)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        db.logp_grad(u, lp, g)
    e1.record(); torch.cuda.synchronize()
    out[f'rw{nw}' if rw else 'cta'] = {'ms': e0.elapsed_time(e1) / 5, 'lp_sum': float(lp.sum()), 'g_abs_sum': float(g.abs().sum())}
    if ref is None:
        ref = (lp.clone(), g.clone())
    else:
        out['max_rel_lp'] = float(((lp - ref[0]).abs() / ref[0].abs()).max())
        out['max_rel_g'] = float(((g - ref[1]).abs() / (ref[1].abs() + 1e-3 * ref[1].abs().max())).max())
print(json.dumps(out))
