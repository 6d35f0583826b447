"""Developer tool: seed-to-seed distribution of the summary table on the reference's fixtures (20 seeds), next to
the reference's published single-run numbers.  usage: python scripts/seed_stats.py [nseeds]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, pandas as pd
from tfcausalimpact_b200 import CausalImpact
GOLD = os.path.join(ROOT, 'tests', 'golden')
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 20


def run(name, make, published):
    rows = []
    for s in range(ns):
        ci = make(s)
        sd = ci.summary_data
        z = 1.959964
        rows.append([sd.loc['predicted', 'average'], (sd.loc['predicted_upper', 'average'] - sd.loc['predicted_lower', 'average']) / (2 * z),
                     sd.loc['predicted_lower', 'average'], sd.loc['predicted_upper', 'average'], ci.p_value])
    a = np.array(rows)
    print(json.dumps({'data': name, 'seeds': ns, 'published': published,
                      'pred_mean': a[:, 0].mean(), 'pred_seed_sd': a[:, 0].std(ddof=1),
                      'sd_mean': a[:, 1].mean(), 'sd_seed_sd': a[:, 1].std(ddof=1),
                      'lower_mean': a[:, 2].mean(), 'lower_seed_sd': a[:, 2].std(ddof=1),
                      'upper_mean': a[:, 3].mean(), 'upper_seed_sd': a[:, 3].std(ddof=1),
                      'p_mean': a[:, 4].mean(), 'p_seed_sd': a[:, 4].std(ddof=1)}), flush=True)


arma = pd.read_csv(os.path.join(GOLD, 'arma_data.csv'))[['y', 'X']]
arma.iloc[70:, 0] += 5
run('arma/vi', lambda s: CausalImpact(arma, [0, 69], [70, 99], model_args={'seed': s}),
    {'README': '120.34 (0.31) [119.76, 120.97] p 0.0', 'R': '120 (0.3)'})
run('arma/hmc', lambda s: CausalImpact(arma, [0, 69], [70, 99], model_args={'seed': s, 'fit_method': 'hmc'}), {})
comp = pd.read_csv(os.path.join(GOLD, 'comparison_data.csv'), index_col=['DATE'])
run('comparison/hmc', lambda s: CausalImpact(comp, ['2019-04-16', '2019-07-14'], ['2019-7-15', '2019-08-01'],
                                            model_args={'fit_method': 'hmc', 'seed': s}),
    {'README': '79282.92 (727.48) [77849.5, 80701.18] p 0.16', 'R': '79232 (736) [77743, 80651] p 0.2'})
run('comparison/vi', lambda s: CausalImpact(comp, ['2019-04-16', '2019-07-14'], ['2019-7-15', '2019-08-01'],
                                           model_args={'seed': s}), {})
goog = pd.read_csv(os.path.join(GOLD, 'google_data.csv'), index_col='date')
run('google/vi', lambda s: CausalImpact(goog, ['2020-01-01', '2020-03-01'], ['2020-03-02', '2020-03-31'], model_args={'seed': s}),
    {'notebook': '120.16 (3.62) [113.72, 127.92] p 0.0', 'R': '129 (4.5) [120, 138]'})
run('google/hmc', lambda s: CausalImpact(goog, ['2020-01-01', '2020-03-01'], ['2020-03-02', '2020-03-31'],
                                        model_args={'seed': s, 'fit_method': 'hmc'}), {})
