"""Pins the ORACLE itself (CPU, no GPU) against results published by the real reference:
README arma table (README.md:61-72: prediction 120.34 (0.31), abs effect 4.89) and the reference's
integer-part known-answer test (tests/test_main.py:299-313), using the oracle's own VI, filter and
forecast restatements end to end (fp64)."""
import os

import numpy as np
import pandas as pd
import torch

from helpers import O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _oracle_causal_impact(y_all, X_all, T, seed):
    """Standardise by pre-period statistics, VI (200 Adam steps, 1 sample), 100 draws, forecast mean."""
    y_pre = y_all[:T]
    mu, sd = y_pre.mean(), y_pre.std()
    xm, xs = X_all[:T].mean(axis=0), X_all[:T].std(axis=0)
    ys = ((y_all - mu) / sd)[None, :]
    Xs = ((X_all - xm) / xs)[None, :, :]
    Tf = len(y_all) - T
    L = O.Layout(T, Tf, X_all.shape[1], 1, 1)
    prob = O.Problem.build(ys[:, :T], Xs, L, dtype=torch.float64)
    mu_q, rho_q, _ = O.vi_fit(prob, np.zeros(1, dtype=np.int64), 200, 1, seed)
    u = O.vi_draw(mu_q, rho_q, 100, seed + 1)[0]                       # [100, P]
    series = torch.zeros(100, dtype=torch.long)
    fm = O.forecast_mean(prob, u, series).mean(dim=0).numpy() * sd + mu   # mixture mean path, original units
    means, vars_ = O.forecast_moments(prob, u, series)
    return fm, y_all[T:], (means.numpy() * sd + mu), vars_.numpy() * sd * sd


def test_oracle_reproduces_readme_arma_table():
    data = pd.read_csv(os.path.join(GOLD, 'arma_data.csv'))[['y', 'X']]
    data.iloc[70:, 0] += 5
    fm, y_post, _, _ = _oracle_causal_impact(data['y'].values.astype(np.float64), data[['X']].values.astype(np.float64),
                                             70, seed=3)
    assert abs(y_post.mean() - 125.23) < 0.01
    assert abs(fm.mean() - 120.34) < 0.3                 # README: 120.34 (0.31)
    assert abs((y_post.mean() - fm.mean()) - 4.89) < 0.3


def test_oracle_reproduces_reference_integer_parts():
    """tests/test_main.py:299-313 of the reference: column 0 (= X) is the response."""
    data = pd.read_csv(os.path.join(GOLD, 'arma_data.csv'))
    data.iloc[70:, 0] += 5
    fm, y_post, _, _ = _oracle_causal_impact(data['X'].values.astype(np.float64), data[['y']].values.astype(np.float64),
                                             70, seed=5)
    assert int(y_post.mean()) == 105
    assert int(fm.mean()) == 100
    assert int(y_post.mean() - fm.mean()) in (4, 5)
