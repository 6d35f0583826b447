import sys; sys.path.insert(0, '.'); sys.path.insert(0, './scripts')
import run_configs as rc
for N in (1250, 2500, 5000):
    rc.run(f'configs[3] slice at {N} series (8 chains x (50 + 100))', N, 365, 30, 50, 7, 'hmc', chains=8, num_results=100, num_warmup=50, predictive_draws=10, niter=500, seed=1)
