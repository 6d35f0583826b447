#!/usr/bin/env python
"""Benchmark of the hot path: fused-on-device HMC transitions (15 leapfrogs, each one Kalman
log-prob + adjoint-gradient evaluation) over a batch of independent series.

    python bench.py --gpus N --steps K --warmup W            # this implementation (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # CPU restatement of the reference path

Metric (BASELINE.json): HMC posterior draws/sec (whole job, all ranks); the Kalman log-prob+grad
evaluation rate and its roofline are reported in the same JSON line.  Workload = BASELINE.json
configs[3]: 10k series x 365 steps, weekly seasonality, SparseLinearRegression with 50 covariates,
8 chains per series.  Headline = weak scaling (every rank owns `--series` series, no data-path collective);
the same JSON line carries a `strong` record -- the NAMED split of configs[3]: `--series` series in TOTAL sharded over
the ranks (`distributed.shard_range`), the [N,21] summary rows gathered over NCCL inside the timed region -- plus a
`full_fit` record (VI initialisation + warm-up + draws + filter + forecast + summary through `causal_impact_batch`),
the configs[4] evaluation and the box's host-to-device copy ceiling.
A "step" is one HMC transition of every chain = N*C posterior draws = N*C*15 evaluations.
"""
import argparse
import ctypes
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

LEAPFROG = 15          # tfp.sts.fit_with_hmc default num_leapfrog_steps (model.py:361-364)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--series', type=int, default=10000, help='series per GPU (weak scaling)')
    ap.add_argument('--chains', type=int, default=8)
    ap.add_argument('--T', type=int, default=365)
    ap.add_argument('--Tf', type=int, default=30)
    ap.add_argument('--k', type=int, default=50)
    ap.add_argument('--nseasons', type=int, default=7)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--init-vi-steps', type=int, default=150, help='untimed VI initialisation (fit_with_hmc: 150)')
    ap.add_argument('--init-vi-samples', type=int, default=5, help='fit_with_hmc: 5')
    ap.add_argument('--init-warmup', type=int, default=50, help='untimed adaptive warm-up transitions (fit_with_hmc: 50)')
    ap.add_argument('--cpu-sample-series', type=int, default=128, help='series of the bounded CPU sample (both CPU legs)')
    ap.add_argument('--no-strong', action='store_true', help='skip the strong-scaling / full-fit records')
    ap.add_argument('--no-extra', action='store_true', help='skip the configs[4] evaluation and copy-ceiling records')
    ap.add_argument('--fit-results', type=int, default=100, help='full_fit: kept draws per chain (fit_with_hmc: 100)')
    ap.add_argument('--fit-warmup', type=int, default=50, help='full_fit: warm-up transitions (fit_with_hmc: 50)')
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# synthetic workload: level random walk + seasonal + sparse regression + noise,
# standardised by the pre-period mean / std as the library does by default
# ------------------------------------------------------------------------------------------------
def synth(N, T, Tf, k, S, seed, device, full=False):
    g = torch.Generator(device=device).manual_seed(seed)
    Tt = T + Tf
    X = torch.randn(N, Tt, max(k, 1), generator=g, device=device)[:, :, :k]
    beta = torch.zeros(N, k, device=device)
    nz = int(math.ceil(k / 5)) if k > 0 else 0
    if nz:
        beta[:, :nz] = torch.randn(N, nz, generator=g, device=device)
    level = torch.cumsum(0.05 * torch.randn(N, Tt, generator=g, device=device), dim=1)
    y = level + 0.3 * torch.randn(N, Tt, generator=g, device=device)
    if S > 1:
        eff = 0.5 * torch.randn(N, S, generator=g, device=device)
        eff = eff - eff.mean(dim=1, keepdim=True)
        idx = torch.arange(Tt, device=device) % S
        y = y + eff[:, idx]
    if k > 0:
        y = y + torch.einsum('ntk,nk->nt', X, beta)
    y[:, T:] += 0.5
    mu = y[:, :T].mean(dim=1, keepdim=True)
    sd = y[:, :T].std(dim=1, unbiased=False, keepdim=True)
    y = (y - mu) / sd
    if full:                                  # + the injected post-period effect in standardised units
        return y[:, :T].contiguous(), y[:, T:].contiguous(), X.contiguous(), (0.5 / sd).reshape(-1)
    return y[:, :T].contiguous(), X.contiguous()


class ClockSampler:
    """nvidia-smi style clock / throttle-reason sampling during the timed region (pynvml)."""

    def __init__(self, index):
        self.samples, self.reasons, self.stop = [], set(), False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
        self.th = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self.nv
        names = {'hw_slowdown': 0x8, 'sw_power_cap': 0x4, 'sw_thermal_slowdown': 0x20,
                 'hw_thermal_slowdown': 0x40, 'hw_power_brake_slowdown': 0x80}
        while not self.stop:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(
                    nv, 'nvmlDeviceGetCurrentClocksEventReasons') else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for n, bit in names.items():
                    if r & bit:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.05)

    def __enter__(self):
        if self.nv is not None:
            self.th.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        if self.nv is not None:
            self.th.join()

    def summary(self):
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': self.max_mhz, 'reasons': ['unavailable']}
        return {'sm_mhz': float(np.median(self.samples)), 'sm_max_mhz': self.max_mhz,
                'reasons': sorted(self.reasons)}


# ------------------------------------------------------------------------------------------------
# CPU restatement of the reference path (oracle port): the same HMC transition, PyTorch fp32 on
# the host cores, dense TFP-style filter + autograd.  TensorFlow Probability itself is not
# installable in this image (no wheel, no network), so `kind` is "port".
# ------------------------------------------------------------------------------------------------
def cpu_setup(args, ns, seed=1):
    from oracle import sts_oracle as O
    y, X = synth(ns, args.T, args.Tf, args.k, args.nseasons, 20244 + seed, 'cpu')
    L = O.Layout(args.T, args.Tf, args.k, args.nseasons, 1)
    prob = O.Problem.build(y.numpy(), X.numpy() if args.k else None, L, dtype=torch.float32)
    B = ns * args.chains
    g = torch.Generator().manual_seed(seed)
    u0 = 0.3 * torch.randn(B, L.P, generator=g)
    u0[:, 0] -= 1.0
    u0[:, 1] -= 2.0
    step0 = torch.full((B, L.P), 0.01)
    series = np.arange(B) // args.chains
    return O, prob, series, u0, step0


def cpu_step_time(args, ns, nsteps=1, seed=1):
    """Seconds per HMC transition on `ns` series x chains (one oracle hmc_run call of `nsteps`
    transitions; its initial-state evaluation is included, 1/(15 nsteps) extra work)."""
    O, prob, series, u0, step0 = cpu_setup(args, ns, seed)
    t0 = time.perf_counter()
    O.hmc_run(prob, series, u0, step0, nsteps, 0, LEAPFROG, seed)
    return (time.perf_counter() - t0) / nsteps


def cpu_sample_series(args):
    """Series of the bounded CPU sample: a CONSTANT (the port is dispatch-bound, so its draws/s depends on the
    sample size -- both CPU legs time the same sample)."""
    return max(1, min(args.cpu_sample_series, args.series))


def tfp_available():
    """The literal reference needs tensorflow + tensorflow_probability (setup.py:51-52); probe site-packages and a
    driver-provided install under baseline/_ref."""
    import importlib.util
    ref = os.path.join(ROOT, 'baseline', '_ref')
    if os.path.isdir(ref) and ref not in sys.path:
        sys.path.insert(0, ref)
    return all(importlib.util.find_spec(m) is not None for m in ('tensorflow', 'tensorflow_probability'))


def run_reference_tfp(args, ns):
    """Literal reference arm: tfp.sts.fit_with_hmc on the model the reference's build_default_model makes
    (causalimpact/model.py:239-321,361-364) with chain_batch_shape=[chains], per series, all host threads.
    Only reachable when TensorFlow Probability is importable (never in this image)."""
    import tensorflow as tf
    import tensorflow_probability as tfp
    y, X = synth(ns, args.T, args.Tf, args.k, args.nseasons, 20245, 'cpu')
    t_total, draws = 0.0, 0
    for i in range(ns):
        obs = tf.convert_to_tensor(y[i].numpy()[:, None], dtype=tf.float32)
        comps = [tfp.sts.LocalLevel(observed_time_series=obs)]
        if args.k:
            comps.append(tfp.sts.SparseLinearRegression(design_matrix=tf.convert_to_tensor(X[i].numpy(), tf.float32),
                                                        weights_prior_scale=0.1))
        if args.nseasons > 1:
            comps.append(tfp.sts.Seasonal(num_seasons=args.nseasons, observed_time_series=obs))
        model = tfp.sts.Sum(comps, observed_time_series=obs)
        t0 = time.perf_counter()
        tfp.sts.fit_with_hmc(model, obs, num_results=args.steps, num_warmup_steps=max(args.warmup, 1),
                             num_leapfrog_steps=LEAPFROG, chain_batch_shape=[args.chains])
        t_total += time.perf_counter() - t0
        draws += args.chains * args.steps
    return draws / t_total, t_total


def run_reference(args, rank, world):
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    ns = cpu_sample_series(args)
    if tfp_available():
        ns = min(ns, 8)
        val, dt = run_reference_tfp(args, ns)
        kind = 'reference'
        sample = (f'{ns} of {args.series} series, tfp.sts.fit_with_hmc(num_results={args.steps}, chain_batch_shape=[{args.chains}]) '
                  f'per series, extrapolated linearly')
    else:
        O, prob, series, u0, step0 = cpu_setup(args, ns)
        O.hmc_run(prob, series, u0, step0, max(args.warmup, 1), 0, LEAPFROG, 3)        # warm-up transitions
        t0 = time.perf_counter()
        O.hmc_run(prob, series, u0, step0, args.steps, 0, LEAPFROG, 4)
        dt = time.perf_counter() - t0
        val = ns * args.chains * args.steps / dt
        kind = 'port'
        sample = (f'{ns} of {args.series} series x {args.chains} chains, {args.steps} HMC transitions x {LEAPFROG} '
                  f'leapfrogs, PyTorch-CPU fp32 dense filter + autograd (tensorflow_probability probed: not importable)')
    line = {
        'impl': 'reference', 'metric': 'hmc_posterior_draws_per_sec', 'value': val, 'unit': 'draws/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args, world),
        'cpu_baseline': {'value': val, 'unit': 'draws/s', 'cores': cores, 'kind': kind, 'sample': sample},
        'e2e': {'value': val, 'unit': 'draws/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'evals_per_sec': val * LEAPFROG, 'gpu_launches': 0,
    }
    print(json.dumps(line))


def workload_config(args, world):
    return {'workload': f'configs[3]: synthetic {args.series} series/GPU x {args.T} steps, nseasons={args.nseasons}, '
                        f'SparseLinearRegression k={args.k}, {args.chains} chains/series, HMC L={LEAPFROG}',
            'series_per_gpu': args.series, 'chains': args.chains, 'T': args.T, 'Tf': args.Tf, 'k': args.k,
            'nseasons': args.nseasons, 'leapfrog': LEAPFROG, 'sharding': f'series x{world} (no collective)',
            'l2': 'inputs (X 0.79 GB + tape) exceed the 126 MB L2; no flush needed'}


# ------------------------------------------------------------------------------------------------
def fit_state(db, C, args, seed):
    """Untimed initialisation exactly as tfp.sts.fit_with_hmc does it: mean-field VI (150 Adam steps x 5
    samples) gives the initial state and per-coordinate step sizes, then 50 warm-up transitions with dual averaging
    on the first 40.  Returns (state u [P,B], frozen step sizes [P,B])."""
    mu, rho = db.vi_fit(C, args.init_vi_samples, args.init_vi_steps, 0.1, 0.01, seed=100 + seed)
    u = db.vi_draw(mu, rho, 1, seed=200 + seed)
    step0 = torch.nn.functional.softplus(rho).contiguous()
    if args.init_warmup > 0:
        _, st0 = db.hmc_run(u, step0, 0, args.init_warmup, int(0.8 * args.init_warmup), LEAPFROG, 0.75, seed=250 + seed)
        step0 = (step0 * st0[2][None, :]).contiguous()           # freeze the adapted step sizes
    db._ws.pop('fit', None)                                      # release the (5x larger) VI workspace
    return u, step0


def max_ranks(x, dist, dev):
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def strong_records(args, rank, world, dev, dist, barrier, weak_ms_per_step):
    """The NAMED configs[3] split: `--series` series in TOTAL, contiguous blocks per rank (distributed.shard_range), all
    chains of a series on one GPU, no data-path collective; the per-series summary rows [n, 21] are all-gathered over
    NCCL inside the timed region.  Also the whole fit end to end (VI-init + warm-up + draws + filter +
    forecast + summary) through batch.causal_impact_batch on the same split."""
    from tfcausalimpact_b200 import distributed as D
    from tfcausalimpact_b200.batch import causal_impact_batch
    from tfcausalimpact_b200.engine import DeviceBatch, reduce_summary
    C, T, Tf, k, S = args.chains, args.T, args.Tf, args.k, args.nseasons
    lo, hi = D.shard_range(args.series, rank, world)
    ns = hi - lo
    y_pre, y_post, X, eff_true = synth(ns, T, Tf, k, S, 30244 + rank, dev, full=True)
    db = DeviceBatch(y_pre, X if k else None, Tf=Tf, nseasons=S)
    B = ns * C
    u, step0 = fit_state(db, C, args, 1000 + rank)
    # summary rows of this shard from the warm state: filter + 200 forecast samples + reduction (K6-K8)
    pm, pv, fs = db.filter_predict(u)
    sims, post_mean = db.forecast_sample(u, fs, 200, 77 + rank, None)
    rows = reduce_summary(y_post, post_mean, sims, 0.05)                        # [ns, 21]
    draws_buf = torch.empty((max(args.steps, args.warmup, 1), db.P, B), dtype=torch.float32, device=dev)
    db.hmc_run(u, step0, max(args.warmup, 1), 0, 0, LEAPFROG, 0.75, seed=311 + rank, draws=draws_buf)
    D.gather_series(rows, args.series)                                          # communicator warm-up
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    db.hmc_run(u, step0, args.steps, 0, 0, LEAPFROG, 0.75, seed=312 + rank, iter_offset=1000, draws=draws_buf)
    allrows = D.gather_series(rows, args.series)
    e1.record()
    barrier()
    ms = max_ranks(e0.elapsed_time(e1), dist, dev)
    gather_ok = bool(allrows.shape[0] == args.series and torch.equal(allrows[lo:hi], rows)
                     and torch.isfinite(allrows).all().item())
    strong = {'scaling': 'strong', 'series_total': args.series, 'series_this_rank': ns, 'value': args.series * C * args.steps / (ms * 1e-3),
              'unit': 'draws/s', 'ms_per_step': ms / args.steps,
              'efficiency_vs_one_gpu': weak_ms_per_step / (world * ms / args.steps),
              'one_gpu_ms_per_step': weak_ms_per_step,
              'collective': f'all_gather of [{args.series},21] summary rows ({"nccl" if world > 1 else "single rank: none"}) inside the timed region',
              'gather_ok': gather_ok, 'finite': bool(torch.isfinite(draws_buf[:args.steps]).all().item())}
    del draws_buf, sims, db
    torch.cuda.empty_cache()
    # ---- the whole fit, end to end, on the same split ----
    barrier()
    t0 = time.perf_counter()
    res = causal_impact_batch(y_pre, y_post, X if k else None, fit_method='hmc', nseasons=S, chains=C,
                              num_results=args.fit_results, num_warmup=args.fit_warmup, predictive_draws=10, niter=500,
                              seed=5 + rank, device=dev)
    summ = D.gather_series(res['summary'].reshape(ns, 20), args.series)
    pval = D.gather_series(res['p_value'].reshape(ns, 1), args.series)
    barrier()
    sec = max_ranks(time.perf_counter() - t0, dist, dev)
    eff = summ[:, 8]                                                            # abs_effect, average column
    full = {'seconds': sec, 'value': args.series * C * args.fit_results / sec, 'unit': 'draws/s',
            'series_per_sec': args.series / sec,
            'what': f'batch.causal_impact_batch on {args.series} series total over {world} GPU(s): VI init 150x5 + '
                    f'{args.fit_warmup} warm-up + {args.fit_results} kept transitions x {C} chains, filter, 500 forecast '
                    f'samples, summary, NCCL gather of summaries',
            'abs_effect_mean': float(eff.mean().item()),
            'abs_effect_true_mean': float(D.gather_series(eff_true.reshape(ns, 1), args.series).mean().item()),
            'accept_rate': float(res['hmc_stats'][0].mean().item()),
            'finite': bool(torch.isfinite(summ).all().item() and torch.isfinite(pval).all().item())}
    return strong, full


def cfg4_eval_record(dev, fp32_peak):
    """configs[4] shape (100 series x 10,000 steps, nseasons=52, no covariates): the few-lane / large-latent evaluation
    (csrc/ci_eval_rw.cu).  A latency-bound kernel: its time follows the SM clock, and this record is taken at the end of a
    long run (7.87 ms at 1965 MHz on a fresh box, ~8.1 ms here when the box sits at 1860-1930 MHz under its power cap)."""
    from tfcausalimpact_b200.engine import DeviceBatch
    N, T, Tf, S = 100, 10000, 520, 52
    y, _ = synth(N, T, Tf, 0, S, 5, dev)
    db = DeviceBatch(y, None, Tf=Tf, nseasons=S)
    u = torch.randn(db.P, N, device=dev) * 0.3
    lp, g = db.logp_grad(u)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        db.logp_grad(u, lp, g)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    d = S
    flops = 3.0 * (5 * d * d + 6 * d + 16) * T * N
    tf = flops / (ms * 1e-3) / 1e12
    return {'workload': 'configs[4]: 100 series x 10,000 steps, nseasons=52, 1 lane per series', 'ms_per_launch': ms,
            'algorithmic_flops_per_launch': flops, 'achieved': tf, 'unit': 'TFLOP/s', 'peak': fp32_peak,
            'frac': tf / fp32_peak, 'finite': bool(torch.isfinite(lp).all().item() and torch.isfinite(g).all().item())}


def h2d_ceiling(dev, dist, barrier, nbytes=256 << 20, reps=4):
    """Plain pinned cudaMemcpyAsync host -> device, one stream per rank, all ranks at once: the box's copy ceiling the
    e2e leg is quoted against."""
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d.copy_(h, non_blocking=True)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    e1.record()
    barrier()
    ms = max_ranks(e0.elapsed_time(e1), dist, dev)
    return nbytes * reps / (ms * 1e-3) / 1e9


def eval_source_sha():
    import hashlib
    hsh = hashlib.sha256()
    for f in ('ci_eval.cu', 'ci_eval_lane.cuh', 'ci_kalman_eff.cuh', 'ci_kalman.cuh', 'ci_params.cuh', 'ci_xstage.cuh', 'ci_core.cuh'):
        hsh.update(open(os.path.join(ROOT, 'tfcausalimpact_b200', 'csrc', f), 'rb').read())
    return hsh.hexdigest()[:16]


# ------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return
    assert torch.cuda.is_available(), 'bench.py (impl b200) needs a CUDA device; there is no CPU fallback'
    torch.cuda.set_device(local)
    dev = torch.device(f'cuda:{local}')
    if world > 1:
        from tfcausalimpact_b200.distributed import bind_to_gpu_numa
        bind_to_gpu_numa(local)          # pinned host buffers of the e2e leg on the GPU's NUMA node
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)

    from tfcausalimpact_b200 import _native as nt
    from tfcausalimpact_b200.engine import DeviceBatch
    lib = nt.lib()
    N, C, T, Tf, k, S = args.series, args.chains, args.T, args.Tf, args.k, args.nseasons
    B = N * C
    y, X = synth(N, T, Tf, k, S, 20244 + rank, dev)
    db = DeviceBatch(y, X if k else None, Tf=Tf, nseasons=S)
    P = db.P
    u, step0 = fit_state(db, C, args, rank)
    draws_buf = torch.empty((max(args.steps, args.warmup, 1), P, B), dtype=torch.float32, device=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: inputs resident in HBM ------------------------------------------------------------
    db.hmc_run(u, step0, max(args.warmup, 1), 0, 0, LEAPFROG, 0.75, seed=301 + rank, draws=draws_buf)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        e0.record()
        _, stats = db.hmc_run(u, step0, args.steps, 0, 0, LEAPFROG, 0.75, seed=302 + rank, iter_offset=1000,
                              draws=draws_buf)
        e1.record()
        barrier()
    ms = e0.elapsed_time(e1)
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms = float(t_ms.item())
    accept = float(stats[0].mean().item())
    finite = bool(torch.isfinite(draws_buf[:args.steps]).all().item())
    draws_per_s = world * B * args.steps / (ms * 1e-3)
    # launches: per transition one boundary kernel (end of the previous + begin of this transition) + L * eval kernel
    # (leapfrog kick/drift fused in); + reset, initial eval, closing boundary kernel, stats
    kern_per_eval = int(lib.ci_eval_num_kernels(ctypes.byref(db.layout(C)))) if hasattr(lib, 'ci_eval_num_kernels') else 5
    launches = args.steps * (1 + LEAPFROG * kern_per_eval) + 3 + kern_per_eval

    # ---- dominant kernel alone: Kalman log-prob + gradient evaluation ------------------------------
    lp = torch.empty(B, dtype=torch.float32, device=dev)
    gr = torch.empty((P, B), dtype=torch.float32, device=dev)
    uu = draws_buf[0].contiguous()
    for _ in range(3):
        db.logp_grad(uu, lp, gr)
    torch.cuda.synchronize()
    n_ev = 20
    e0.record()
    for _ in range(n_ev):
        db.logp_grad(uu, lp, gr)
    e1.record()
    torch.cuda.synchronize()
    eval_ms = e0.elapsed_time(e1) / n_ev
    d = 1 + (S - 1 if S > 1 else 0)
    w_eval = 3 * (5 * d * d + 6 * d + 2 * k + 16)                      # flops / step / pair
    flops_per_eval = float(w_eval) * T * B
    bytes_per_eval = float(N) * (4 * T * (k + 1) + C * (8 * P + 4))    # algorithmic bytes
    # measured FP32-FMA peak (MEASURED_PEAKS.json has HBM and bf16 tensor figures only)
    probe = torch.zeros(1, device=dev)
    fl = ctypes.c_double(0.0)
    for _ in range(2):
        nt.check(lib.ci_bench_fma(4096, 148 * 8, nt.ptr(probe), nt.stream_ptr(), ctypes.byref(fl)), 'ci_bench_fma')
    torch.cuda.synchronize()
    e0.record()
    nt.check(lib.ci_bench_fma(4096, 148 * 8, nt.ptr(probe), nt.stream_ptr(), ctypes.byref(fl)), 'ci_bench_fma')
    e1.record()
    torch.cuda.synchronize()
    fp32_peak = fl.value / (e0.elapsed_time(e1) * 1e-3) / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, 'profiles', 'r2_eval_traffic.json')))
        wl = tj['workload']
        # dram read + write of one launch from ONE `ncu --set full` capture (scripts/ncu_traffic.py); refused when the
        # kernel sources changed since the capture or the workload differs
        if (wl['series'], wl['chains'], wl['T'], wl['k'], wl['nseasons']) == (N, C, T, k, S) and \
                tj.get('source_sha') == eval_source_sha():
            traffic = tj['traffic_bytes_per_launch']
    except Exception:
        pass
    ach_tf = flops_per_eval / (eval_ms * 1e-3) / 1e12
    ach_gbs = bytes_per_eval / (eval_ms * 1e-3) / 1e9
    roofline = {
        'bound': 'fp32', 'kernel': 'ci_kalman_logp_grad (K1-K3 evaluation)', 'achieved': ach_tf,
        'peak': fp32_peak, 'unit': 'TFLOP/s', 'frac': ach_tf / fp32_peak, 'traffic': traffic,
        'peak_source': 'measured in this run (ci_bench_fma FFMA probe); nominal 148 SM x 128 x 2 x 1.965 GHz = 74.4',
        'frac_of_nominal': ach_tf / 74.4, 'ms_per_launch': eval_ms,
        'algorithmic_flops_per_launch': flops_per_eval, 'algorithmic_bytes_per_launch': bytes_per_eval,
        'hbm': {'bound': 'hbm', 'achieved': ach_gbs, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': ach_gbs / hbm_peak,
                'traffic': traffic, 'actual_traffic_frac': (traffic / (eval_ms * 1e-3) / 1e9 / hbm_peak) if traffic else None,
                'peak_source': 'measured (MEASURED_PEAKS.json)' if peaks else 'fallback'},
    }

    # ---- e2e: host buffers in, draws out, through the public engine API ----------------------------
    e2e = None
    if not args.no_e2e:
        y_h = y.cpu().pin_memory()
        X_h = X.cpu().pin_memory() if k else None
        u_h = u.cpu().pin_memory()
        s_h = step0.cpu().pin_memory()
        out_h = torch.empty((P, B), dtype=torch.float32).pin_memory()
        h2d = y_h.numel() * 4 + (X_h.numel() * 4 if k else 0) + u_h.numel() * 4 + s_h.numel() * 4
        d2h = out_h.numel() * 4

        from tfcausalimpact_b200.engine import HostUploader
        up = HostUploader(dev)

        def e2e_compute(staged, seed):
            y_d, X_d, ue, se = HostUploader.finish(*staged)
            dbe = DeviceBatch(y_d, X_d, Tf=Tf, nseasons=S, device=dev)
            dr, _ = dbe.hmc_run(ue, se, 1, 0, 0, LEAPFROG, 0.75, seed=seed, draws=draws_buf[:1])
            out_h.copy_(dr[0], non_blocking=False)

        # every step uploads ITS OWN inputs (y, X, state, step sizes) from pinned host memory; the
        # upload of step i+1 runs on a side stream while step i computes (double buffering)
        staged = up.start(y_h, X_h, u_h, s_h)
        for i in range(min(args.warmup, 3)):
            nxt = up.start(y_h, X_h, u_h, s_h)
            e2e_compute(staged, 400 + i)
            staged = nxt
        barrier()
        n_e2e = max(1, min(args.steps, 10))
        e0.record()
        for i in range(n_e2e):
            nxt = up.start(y_h, X_h, u_h, s_h)      # K uploads inside the timed region (the last one is the next step's)
            e2e_compute(staged, 500 + i)
            staged = nxt
        e1.record()
        barrier()
        ms_e = e0.elapsed_time(e1)
        t_e = torch.tensor([ms_e], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
        ms_e = float(t_e.item())
        ceil_gbs = h2d_ceiling(dev, dist, barrier)
        e2e = {'value': world * B * n_e2e / (ms_e * 1e-3), 'unit': 'draws/s', 'h2d_bytes_per_step': int(h2d),
               'd2h_bytes_per_step': int(d2h), 'steps': n_e2e, 'ms_per_step': ms_e / n_e2e,
               'h2d_gbs_per_rank': h2d / (ms_e / n_e2e * 1e-3) / 1e9,
               'h2d_ceiling_gbs_per_rank': ceil_gbs,       # plain pinned cudaMemcpyAsync, all ranks at once
               'h2d_frac_of_ceiling': h2d / (ms_e / n_e2e * 1e-3) / 1e9 / ceil_gbs,
               'api': 'engine.HostUploader (pinned host y, X, state; double-buffered) -> DeviceBatch -> hmc_run(1 transition) -> draws to pinned host'}

    # ---- the named split (strong scaling) + the whole fit; configs[4] evaluation ----------------------
    del draws_buf, lp, gr, uu
    strong = full_fit = cfg4 = None
    if not args.no_strong:
        if world > 1:
            del db
        torch.cuda.empty_cache()
        strong, full_fit = strong_records(args, rank, world, dev, dist, barrier, ms / args.steps)
    if rank == 0 and not args.no_extra:
        cfg4 = cfg4_eval_record(dev, fp32_peak)
    if dist is not None:
        barrier()

    # ---- CPU baseline (rank 0 only, N=1 only) -------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count() or 1)
        ns = cpu_sample_series(args)
        cpu_step_time(args, ns, 1, seed=8)            # warm-up transition (thread pool, allocator)
        t = cpu_step_time(args, ns, 2, seed=9)
        cpu = {'value': ns * C / t, 'unit': 'draws/s', 'cores': torch.get_num_threads(), 'kind': 'port',
               'sample': f'{ns} of {N} series x {C} chains, 2 HMC transitions x {LEAPFROG} leapfrogs, PyTorch-CPU '
                         f'fp32 dense filter + autograd (oracle port; TFP not installable here)'}

    if rank == 0:
        line = {
            'metric': 'hmc_posterior_draws_per_sec', 'value': draws_per_s, 'unit': 'draws/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(args, world), 'clocks': clk.summary(), 'e2e': e2e, 'gpu_launches': launches,
            'roofline': roofline, 'cpu_baseline': cpu, 'strong': strong, 'full_fit': full_fit, 'configs4_eval': cfg4,
            'evals_per_sec': draws_per_s * LEAPFROG, 'eval_only_evals_per_sec_per_gpu': B / (eval_ms * 1e-3),
            'accept_rate': accept, 'finite': finite,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
