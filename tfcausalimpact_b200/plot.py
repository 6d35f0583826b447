"""Three-panel figure of the inference frame (presentation only; matplotlib imported lazily, as
the reference does at causalimpact/plot.py:170-179)."""
from typing import List, Tuple

import numpy as np
import pandas as pd


def get_plotter():  # pragma: no cover
    import matplotlib.pyplot as plt
    return plt


def plot(inferences: pd.DataFrame, pre_data: pd.DataFrame, post_data: pd.DataFrame,
         panels: List[str] = ['original', 'pointwise', 'cumulative'], figsize: Tuple[int] = (10, 7),
         show: bool = True) -> None:
    plt = get_plotter()
    valid = {'original', 'pointwise', 'cumulative'}
    for panel in panels:
        if panel not in valid:
            raise ValueError(f'"{panel}" is not a valid panel. Valid panels are: '
                             '"original", "pointwise", "cumulative".')
    fig = plt.figure(figsize=figsize)
    pre_post_index = pre_data.index.union(post_data.index)
    # first point is skipped: the diffuse initial state makes its interval uninformative
    inferences = inferences.iloc[1:, :]
    intervention_idx = pre_post_index.get_loc(post_data.index[0])
    n = len(panels)
    ax = None
    pos = 1
    if 'original' in panels:
        ax = plt.subplot(n, 1, pos); pos += 1
        y = pd.concat([pre_data.iloc[1:, 0], post_data.iloc[:, 0]])
        ax.plot(y, 'k', label='y')
        ax.plot(inferences['complete_preds_means'], 'orangered', ls='dashed', label='Predicted')
        ax.axvline(pre_post_index[intervention_idx - 1], c='gray', linestyle='--')
        ax.fill_between(inferences['complete_preds_means'].index, inferences['complete_preds_lower'],
                        inferences['complete_preds_upper'], color=(1.0, 0.4981, 0.0549), alpha=0.4)
        ax.legend()
        ax.grid(True, color='gainsboro')
    if 'pointwise' in panels:
        ax = plt.subplot(n, 1, pos, sharex=ax); pos += 1
        ax.plot(inferences['point_effects_means'], ls='dashed', color='orangered', label='Point Effects')
        ax.axvline(pre_post_index[intervention_idx - 1], c='gray', linestyle='--')
        ax.fill_between(inferences['point_effects_means'].index, inferences['point_effects_lower'],
                        inferences['point_effects_upper'], color=(1.0, 0.4981, 0.0549), alpha=0.4)
        ax.axhline(y=0, color='gray')
        ax.legend()
        ax.grid(True, color='gainsboro')
    if 'cumulative' in panels:
        ax = plt.subplot(n, 1, pos, sharex=ax)
        ax.plot(inferences['post_cum_effects_means'], ls='dashed', color='orangered', label='Cumulative Effect')
        ax.axvline(pre_post_index[intervention_idx - 1], c='gray', linestyle='--')
        ax.fill_between(inferences['post_cum_effects_means'].index, inferences['post_cum_effects_lower'],
                        inferences['post_cum_effects_upper'], color=(1.0, 0.4981, 0.0549), alpha=0.4)
        ax.axhline(y=0, color='gray', linestyle='--')
        ax.legend()
        ax.grid(True, color='gainsboro')
    if show:
        plt.show()
