"""Three-panel figure of the inference frame (presentation only, host side; matplotlib is imported lazily through
`get_plotter`, which tests replace by a mock exactly as the reference's do -- causalimpact/plot.py:25-179).

The matplotlib call sequence per panel (x = the union of the pre / post index without its first point, whose diffuse
initial state makes the interval uninformative) is the reference's, so a figure produced here is the same figure:
observed series in black (original panel only), dashed orange-red mean, grey dashed line at the last pre-period point,
translucent band between the lower / upper columns, zero line on the effect panels, legend and light grid."""
from typing import List, Tuple

import numpy as np
import pandas as pd

_VALID = ('original', 'pointwise', 'cumulative')
_BAND = (1.0, 0.4981, 0.0549)
# panel -> (column prefix of the inference frame, legend label, zero-line keyword arguments or None)
_PANELS = {
    'original': ('complete_preds', 'Predicted', None),
    'pointwise': ('point_effects', 'Point Effects', {'color': 'gray'}),
    'cumulative': ('post_cum_effects', 'Cumulative Effect', {'color': 'gray', 'linestyle': '--'}),
}


def get_plotter():  # pragma: no cover
    import matplotlib.pyplot as plt
    return plt


def build_data(pre_data: pd.DataFrame, post_data: pd.DataFrame, inferences: pd.DataFrame):
    """Drops the pre-period points whose response is NaN (gaps that regularisation inserted, data.py) from the data and
    from the inference frame, and casts everything to float64.  With a positional inference index the data frames are
    re-indexed positionally first (reference behaviour, causalimpact/plot.py:148-167)."""
    if isinstance(inferences.index, pd.RangeIndex):
        n_pre, n_post = len(pre_data), len(post_data)
        pre_data = pre_data.set_index(pd.RangeIndex(0, n_pre))
        post_data = post_data.set_index(pd.RangeIndex(n_pre, n_pre + n_post))
    holes = pre_data.index[pre_data.iloc[:, 0].isnull()]
    return (pre_data.drop(holes).astype(np.float64), post_data.astype(np.float64),
            inferences.drop(holes).astype(np.float64))


def plot(inferences: pd.DataFrame, pre_data: pd.DataFrame, post_data: pd.DataFrame,
         panels: List[str] = ['original', 'pointwise', 'cumulative'], figsize: Tuple[int] = (10, 7),
         show: bool = True) -> None:
    plt = get_plotter()
    plt.figure(figsize=figsize)
    for panel in panels:
        if panel not in _VALID:
            raise ValueError('"{}" is not a valid panel. Valid panels are: {}.'.format(
                panel, ', '.join('"{}"'.format(v) for v in _VALID)))
    pre_data, post_data, inferences = build_data(pre_data, post_data, inferences)
    index = pre_data.index.union(post_data.index)
    last_pre = index[index.get_loc(post_data.index[0]) - 1]
    x = index[1:]
    wanted = [p for p in _VALID if p in panels]
    ax = plt.subplot(len(panels), 1, 1)
    for pos, panel in enumerate(wanted, start=1):
        prefix, label, zero_line = _PANELS[panel]
        if panel == 'original':
            ax.plot(index, pd.concat([pre_data.iloc[:, 0], post_data.iloc[:, 0]]), 'k', label='y')
            ax.plot(x, inferences[prefix + '_means'].iloc[1:], color='orangered', ls='dashed', label=label)
        else:
            ax = plt.subplot(len(panels), 1, pos, sharex=ax)
            ax.plot(x, inferences[prefix + '_means'].iloc[1:], ls='dashed', color='orangered', label=label)
        ax.axvline(last_pre, c='gray', linestyle='--')
        ax.fill_between(x, inferences[prefix + '_lower'].iloc[1:], inferences[prefix + '_upper'].iloc[1:],
                        color=_BAND, alpha=0.4)
        if zero_line is not None:
            ax.axhline(y=0, **zero_line)
        ax.legend()
        ax.grid(True, color='gainsboro')
        if pos != len(panels) and panel != 'cumulative':
            plt.setp(ax.get_xticklabels(), visible=False)
    if show:
        plt.show()
