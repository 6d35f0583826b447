"""plot.py against a mocked matplotlib, in the style of the reference's tests/test_plot.py (which monkeypatches
`causalimpact.plot.get_plotter`): the exact call sequence per panel, the NaN handling of `build_data`, index kinds
(positional / dates / gaps) and the invalid-panel error.  Host only, no GPU."""
from unittest import mock

import numpy as np
import pandas as pd
import pytest
from numpy.testing import assert_array_equal

import tfcausalimpact_b200.plot as plotter

BAND = (1.0, 0.4981, 0.0549)
COLS = ['complete_preds_means', 'complete_preds_lower', 'complete_preds_upper',
        'point_effects_means', 'point_effects_lower', 'point_effects_upper',
        'post_cum_effects_means', 'post_cum_effects_lower', 'post_cum_effects_upper']


def _frames(n=200, n_pre=100, dates=False, gap=0):
    rng = np.random.default_rng(7)
    data = pd.DataFrame(rng.normal(size=(n, 3)), columns=['y', 'x1', 'x2'])
    inf = pd.DataFrame(rng.random(size=(n, 9)), columns=COLS)
    if dates:
        data.index = pd.date_range('2020-01-01', periods=n)
        inf.index = data.index
    pre, post = data.iloc[:n_pre], data.iloc[n_pre + gap:]
    return inf, pre, post


def _run(monkeypatch, inf, pre, post, **kw):
    plt = mock.Mock()
    monkeypatch.setattr(plotter, 'get_plotter', plt)
    plotter.plot(inf, pre, post, **kw)
    return plt


def test_build_data_drops_nan_pre_points():
    pre = pd.DataFrame([0, 1, np.nan])
    post = pd.DataFrame([3, 4], index=[3, 4])
    inf = pd.DataFrame([0, 1, 2, 3, 4])
    pre2, post2, inf2 = plotter.build_data(pre, post, inf)
    pd.testing.assert_frame_equal(pre2, pd.DataFrame([0, 1]).astype(np.float64))
    pd.testing.assert_frame_equal(post2, pd.DataFrame([3, 4], index=[3, 4]).astype(np.float64))
    pd.testing.assert_frame_equal(inf2, pd.DataFrame([0, 1, 3, 4], index=[0, 1, 3, 4]).astype(np.float64))


@pytest.mark.parametrize('dates,gap', [(False, 0), (True, 0), (False, 10), (True, 10)])
def test_original_panel_calls(monkeypatch, dates, gap):
    inf, pre, post = _frames(dates=dates, gap=gap)
    if gap and not dates:
        inf = inf.iloc[:len(pre) + len(post)]
    if gap and dates:
        inf = inf.drop(inf.index[len(pre):len(pre) + gap])
    plt = _run(monkeypatch, inf, pre, post, panels=['original'])
    pre_b, post_b, inf_b = plotter.build_data(pre, post, inf)
    idx = pre_b.index.union(post_b.index)
    plt.assert_called_once()
    plt.return_value.figure.assert_called_with(figsize=(10, 7))
    plt.return_value.subplot.assert_any_call(1, 1, 1)
    ax = plt.return_value.subplot.return_value
    calls = ax.plot.call_args_list
    assert_array_equal(calls[0][0][0], idx)
    assert_array_equal(calls[0][0][1], pd.concat([pre_b.iloc[:, 0], post_b.iloc[:, 0]]))
    assert calls[0][0][2] == 'k' and calls[0][1] == {'label': 'y'}
    assert_array_equal(calls[1][0][0], idx[1:])
    assert_array_equal(calls[1][0][1], inf_b['complete_preds_means'].iloc[1:])
    assert calls[1][1] == {'color': 'orangered', 'ls': 'dashed', 'label': 'Predicted'}
    ax.axvline.assert_called_with(pre_b.index[-1], c='gray', linestyle='--')
    fb = ax.fill_between.call_args_list[0]
    assert_array_equal(fb[0][0], idx[1:])
    assert_array_equal(fb[0][1], inf_b['complete_preds_lower'].iloc[1:])
    assert_array_equal(fb[0][2], inf_b['complete_preds_upper'].iloc[1:])
    assert fb[1] == {'color': BAND, 'alpha': 0.4}
    ax.grid.assert_called_with(True, color='gainsboro')
    ax.legend.assert_called()
    ax.axhline.assert_not_called()
    plt.return_value.show.assert_called_once()


@pytest.mark.parametrize('panel,prefix,label,zero_kw', [
    ('pointwise', 'point_effects', 'Point Effects', {'y': 0, 'color': 'gray'}),
    ('cumulative', 'post_cum_effects', 'Cumulative Effect', {'y': 0, 'color': 'gray', 'linestyle': '--'}),
])
def test_effect_panel_calls(monkeypatch, panel, prefix, label, zero_kw):
    inf, pre, post = _frames()
    plt = _run(monkeypatch, inf, pre, post, panels=[panel])
    idx = pre.index.union(post.index)
    first_ax = plt.return_value.subplot.return_value
    plt.return_value.subplot.assert_any_call(1, 1, 1, sharex=first_ax)
    ax = first_ax
    call = ax.plot.call_args_list[0]
    assert_array_equal(call[0][0], idx[1:])
    assert_array_equal(call[0][1], inf[prefix + '_means'].iloc[1:])
    assert call[1] == {'ls': 'dashed', 'color': 'orangered', 'label': label}
    ax.axvline.assert_called_with(pre.index[-1], c='gray', linestyle='--')
    fb = ax.fill_between.call_args_list[0]
    assert_array_equal(fb[0][1], inf[prefix + '_lower'].iloc[1:])
    assert_array_equal(fb[0][2], inf[prefix + '_upper'].iloc[1:])
    assert fb[1] == {'color': BAND, 'alpha': 0.4}
    ax.axhline.assert_called_with(**zero_kw)
    plt.return_value.show.assert_called_once()


def test_three_panels_share_the_axis_and_hide_inner_tick_labels(monkeypatch):
    inf, pre, post = _frames()
    plt = _run(monkeypatch, inf, pre, post, show=False)
    sub = plt.return_value.subplot
    ax = sub.return_value
    assert [c[0] for c in sub.call_args_list] == [(3, 1, 1), (3, 1, 2), (3, 1, 3)]
    assert sub.call_args_list[1][1] == {'sharex': ax} and sub.call_args_list[2][1] == {'sharex': ax}
    assert ax.plot.call_count == 4 and ax.fill_between.call_count == 3 and ax.axhline.call_count == 2
    assert plt.return_value.setp.call_count == 2        # x tick labels of the two upper panels
    plt.return_value.show.assert_not_called()


def test_nan_in_pre_data_is_removed_before_plotting(monkeypatch):
    inf, pre, post = _frames()
    pre = pre.copy()
    pre.iloc[5:8, 0] = np.nan
    plt = _run(monkeypatch, inf, pre, post, panels=['original'])
    ax = plt.return_value.subplot.return_value
    x0, y0 = ax.plot.call_args_list[0][0][:2]
    assert len(x0) == len(y0) == 197 and not np.isnan(np.asarray(y0, dtype=float)).any()
    assert 5 not in list(x0) and 7 not in list(x0)


def test_invalid_panel_raises_the_reference_message(monkeypatch):
    inf, pre, post = _frames()
    with pytest.raises(ValueError) as exc:
        _run(monkeypatch, inf, pre, post, panels=['original', 'test'])
    assert str(exc.value) == ('"test" is not a valid panel. Valid panels are: "original", "pointwise", "cumulative".')
