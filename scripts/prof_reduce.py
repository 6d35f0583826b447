"""Profile target: ci_reduce_inferences on 2000 series x 1000 draws x (365 + 30) steps."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tfcausalimpact_b200.engine import reduce_inferences
N, T, Tf, ns = 2000, 365, 30, 1000
g = torch.Generator(device='cuda').manual_seed(0)
y = torch.randn(N, T, generator=g, device='cuda'); yp = torch.randn(N, Tf, generator=g, device='cuda')
pre = torch.randn(N, ns, T, generator=g, device='cuda'); post = torch.randn(N, ns, Tf, generator=g, device='cuda')
for _ in range(2):
    out = reduce_inferences(y, yp, pre, pre.mean(1), post, post.mean(1), 0.05)
torch.cuda.synchronize()
print('ok', bool(torch.isfinite(out).all()))
