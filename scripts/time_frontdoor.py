"""Wall time of the single-series front door on the reference's own examples (configs[0], configs[1])."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, pandas as pd, torch
from tfcausalimpact_b200 import CausalImpact
GOLD = os.path.join(ROOT, 'tests', 'golden')
arma = pd.read_csv(os.path.join(GOLD, 'arma_data.csv'))[['y', 'X']]
arma.iloc[70:, 0] += 5
comp = pd.read_csv(os.path.join(GOLD, 'comparison_data.csv'), index_col=['DATE'])
for rep in range(3):
    np.random.seed(rep)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ci = CausalImpact(arma, [0, 69], [70, 99])
    torch.cuda.synchronize(); t1 = time.perf_counter()
    ci2 = CausalImpact(comp, ['2019-04-16', '2019-07-14'], ['2019-7-15', '2019-08-01'], model_args={'fit_method': 'hmc'})
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(json.dumps({'rep': rep, 'arma_vi_seconds': t1 - t0, 'comparison_hmc_seconds': t2 - t1,
                      'arma_pred': float(ci.summary_data.loc['predicted', 'average']),
                      'comparison_pred': float(ci2.summary_data.loc['predicted', 'average']),
                      'comparison_p': float(ci2.p_value), 'accept': ci2.model_kernel_results['accept_rate']}), flush=True)
print(ci.summary())
print(ci2.summary())
