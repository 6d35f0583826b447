"""End-to-end runs of BASELINE.json configs[2] and configs[4] (and a slice of configs[3]) through the
batched front door: fit -> filter -> forecast samples -> summary, with timings and effect recovery."""
import json, math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from tfcausalimpact_b200.batch import causal_impact_batch


def run(name, N, T, Tf, k, S, method, **kw):
    y_all, X = bench.synth(N, T, Tf, k, S, 77, 'cuda')            # standardised pre-period, post += 0.5 (pre-sd units)
    # rebuild post-period values: synth returns only y[:, :T]; regenerate the full series deterministically
    g = torch.Generator(device='cuda').manual_seed(77)
    Tt = T + Tf
    Xf = torch.randn(N, Tt, max(k, 1), generator=g, device='cuda')[:, :, :k]
    beta = torch.zeros(N, k, device='cuda')
    nz = int(math.ceil(k / 5)) if k > 0 else 0
    if nz:
        beta[:, :nz] = torch.randn(N, nz, generator=g, device='cuda')
    level = torch.cumsum(0.05 * torch.randn(N, Tt, generator=g, device='cuda'), dim=1)
    y = level + 0.3 * torch.randn(N, Tt, generator=g, device='cuda')
    if S > 1:
        eff = 0.5 * torch.randn(N, S, generator=g, device='cuda')
        eff = eff - eff.mean(dim=1, keepdim=True)
        y = y + eff[:, torch.arange(Tt, device='cuda') % S]
    if k > 0:
        y = y + torch.einsum('ntk,nk->nt', Xf, beta)
    y[:, T:] += 0.5
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = causal_impact_batch(y[:, :T].contiguous(), y[:, T:].contiguous(), Xf if k else None, fit_method=method,
                              nseasons=S, **kw)
    summ = res['summary']
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    eff_hat = summ[:, 4, 0]
    out = {'config': name, 'N': N, 'T': T, 'Tf': Tf, 'k': k, 'nseasons': S, 'fit_method': method, 'seconds': dt,
           'series_per_sec': N / dt, 'abs_effect_mean': float(eff_hat.mean()), 'abs_effect_true': 0.5,
           'abs_effect_rmse': float(((eff_hat - 0.5) ** 2).mean().sqrt()),
           'frac_p_below_0.05': float((res['p_value'] < 0.05).float().mean()),
           'finite': bool(torch.isfinite(summ).all())}
    if res['hmc_stats'] is not None:
        out['accept_rate'] = float(res['hmc_stats'][0].mean())
    out.update({k_: v for k_, v in kw.items() if isinstance(v, (int, float))})
    print(json.dumps(out), flush=True)


if __name__ == '__main__':
    which = sys.argv[1:] or ['2', '4']
    if '2' in which:
        run('configs[2]: 1k series x 365, nseasons=7, 10 covariates, HMC 8 chains x 1000 draws', 1000, 365, 30, 10, 7,
            'hmc', chains=8, num_results=1000, num_warmup=500, predictive_draws=25, niter=1000, seed=1)
    if '3' in which:
        run('configs[3] slice: 10k series x 365, nseasons=7, 50 covariates, HMC 8 chains x 100 draws', 10000, 365, 30,
            50, 7, 'hmc', chains=8, num_results=100, num_warmup=50, predictive_draws=10, niter=500, seed=1)
    if 'd' in which:
        # the reference's DEFAULT settings at scale: one chain / one VI sample per series (model.py:361-386)
        run('defaults, VI: 10k series x 365, nseasons=7, 50 covariates, fit_method=vi (200 steps), 1000 samples', 10000,
            365, 30, 50, 7, 'vi', niter=1000, seed=1)
        run('defaults, HMC: 10k series x 365, nseasons=7, 50 covariates, 1 chain x (50 warm-up + 100 draws)', 10000,
            365, 30, 50, 7, 'hmc', chains=1, num_results=100, num_warmup=50, niter=1000, seed=1)
    if '4' in which:
        run('configs[4]: 100 series x 10k steps, nseasons=52, VI + 1000-sample forecast', 100, 10000, 520, 0, 52, 'vi',
            niter=1000, seed=1)
