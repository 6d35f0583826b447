"""Host-side logic of the front door (no GPU): argument validation with the reference's error
strings (/root/reference/tests/test_model.py:28-95, tests/test_data.py), period slicing,
standardisation, summary rendering against the reference's golden files, model spec layout."""
import os

import numpy as np
import pandas as pd
import pytest

from tfcausalimpact_b200 import data as cidata
from tfcausalimpact_b200 import misc
from tfcausalimpact_b200 import model as cimodel
from tfcausalimpact_b200 import summary as summarizer

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


@pytest.fixture
def rand_data():
    return pd.DataFrame(np.random.default_rng(0).normal(size=(200, 3)), columns=['y', 'x1', 'x2'])


@pytest.fixture
def summary_data():
    data = [[5.123, 10.456], [4.234, 9.567], [3.123, 8.456], [6.234, 9.567], [3.123, 10.456],
            [2.234, 4.567], [6.123, 9.456], [0.2345, 0.5678], [0.1234, 0.4567], [0.2345, 0.5678]]
    return pd.DataFrame(data, columns=['average', 'cumulative'],
                        index=['actual', 'predicted', 'predicted_lower', 'predicted_upper', 'abs_effect',
                               'abs_effect_lower', 'abs_effect_upper', 'rel_effect', 'rel_effect_lower',
                               'rel_effect_upper'])


def test_process_model_args_defaults_and_errors():
    out = cimodel.process_model_args({})
    assert out == {'standardize': True, 'prior_level_sd': 0.01, 'niter': 1000, 'fit_method': 'vi',
                   'nseasons': 1, 'season_duration': 1}
    for args, msg in [({'standardize': 'yes'}, 'standardize argument must be of type bool.'),
                      ({'prior_level_sd': 1}, 'prior_level_sd argument must be of type float.'),
                      ({'niter': 10.5}, 'niter argument must be of type int.'),
                      ({'fit_method': 'test'}, 'fit_method can be either "hmc" or "vi".'),
                      ({'nseasons': 'test'}, 'nseasons argument must be of type int.'),
                      ({'season_duration': 'test'}, 'season_duration argument must be of type int.'),
                      ({'season_duration': 2}, 'nseasons must be bigger than 1 when season_duration is also '
                                               'bigger than 1.')]:
        with pytest.raises(ValueError) as e:
            cimodel.process_model_args(args)
        assert str(e.value) == msg


def test_period_and_data_validation(rand_data):
    with pytest.raises(ValueError) as e:
        cidata.process_input_data(None, [0, 10], [11, 20], None, {}, 0.05)
    assert str(e.value) == 'data input argument cannot be empty'
    with pytest.raises(ValueError) as e:
        cidata.process_input_data(None, None, [11, 20], None, {}, 0.05)
    assert str(e.value) == 'data, pre_period input arguments cannot be empty'
    with pytest.raises(ValueError) as e:
        cidata.process_pre_post_data(rand_data, [5, 10], [9, 20])
    assert str(e.value).startswith('Values in training data cannot be present')
    with pytest.raises(ValueError) as e:
        cidata.process_pre_post_data(rand_data, [5, 7], [8, 20])
    assert str(e.value) == 'pre_period must span at least 3 time points.'
    with pytest.raises(ValueError) as e:
        cidata.process_pre_post_data(rand_data, [0, 100], [100, 120])
    assert str(e.value) == 'post_period first value (100) must be bigger than the second value of pre_period (100).'
    with pytest.raises(ValueError) as e:
        cidata.process_period((0, 10), rand_data)
    assert str(e.value) == 'Input period must be of type list.'
    with pytest.raises(ValueError) as e:
        cidata.process_period([0, 500], rand_data)
    assert str(e.value) == '500 not present in input data index.'
    with pytest.raises(ValueError) as e:
        cidata.process_period([0.5, 1.5], rand_data)
    assert str(e.value) == 'Input periods must contain either int, str or pandas Timestamp'
    with pytest.raises(ValueError) as e:
        cidata.process_alpha(1)
    assert str(e.value) == 'alpha must be of type float.'
    with pytest.raises(ValueError) as e:
        cidata.process_alpha(1.5)
    assert str(e.value) == 'alpha must range between 0 (zero) and 1 (one) inclusive.'
    bad = rand_data.copy()
    bad.iloc[3, 1] = np.nan
    with pytest.raises(ValueError) as e:
        cidata.format_input_data(bad)
    assert str(e.value) == 'Input data cannot have NAN values.'
    const = rand_data.copy()
    const['y'] = 1.0
    with pytest.raises(ValueError) as e:
        cidata.format_input_data(const)
    assert str(e.value) == 'Input response cannot be constant.'


def test_pre_post_slicing_range_and_dates(rand_data):
    pre, post = cidata.process_pre_post_data(cidata.format_input_data(rand_data), [0, 99], [100, 199])
    assert len(pre) == 100 and len(post) == 100
    assert isinstance(pre.index, pd.DatetimeIndex) and pre.index[0] == pd.Timestamp('2020-01-01')
    assert pre.dtypes.iloc[0] == np.float32
    dated = rand_data.set_index(pd.date_range('2018-01-01', periods=200))
    pre, post = cidata.process_pre_post_data(dated, ['20180101', '20180410'], ['20180411', '20180719'])
    assert len(pre) == 100 and len(post) == 100


def test_regularize_matches_asfreq_and_masks_gaps():
    """tests/test_data.py:52-58 of the reference: regularisation == asfreq(1 day)."""
    btc = pd.read_csv(os.path.join(GOLD, 'btc.csv'), parse_dates=True, index_col='Date')
    reg = cidata.regularize_series(btc)
    pd.testing.assert_frame_equal(reg, btc.asfreq(pd.DateOffset(days=1)))
    assert reg.iloc[:, 0].isna().sum() > 0
    mask = cidata._build_nan_mask(reg)
    assert mask.sum() == len(btc)


def test_standardize_roundtrip():
    df = pd.DataFrame(np.random.default_rng(1).normal(3, 2, size=(50, 2)))
    normed, (mu, sig) = misc.standardize(df)
    np.testing.assert_allclose(normed.mean().values, 0, atol=1e-12)
    np.testing.assert_allclose(normed.std(ddof=0).values, 1, atol=1e-12)
    np.testing.assert_allclose(misc.unstandardize(normed, (mu, sig)).values, df.values, atol=1e-12)
    with pytest.raises(ValueError):
        misc.standardize(df.iloc[:1])
    assert misc.get_z_score(0.5) == 0.0
    assert round(misc.get_z_score(0.9177), 2) == 1.39
    assert misc.maybe_unstandardize(df, None) is df


def test_summary_matches_reference_goldens(summary_data):
    exp = open(os.path.join(GOLD, 'test_output_summary_1')).read().strip()
    assert summarizer.summary(summary_data, 0.459329, alpha=0.1) == exp
    exp = open(os.path.join(GOLD, 'test_output_summary_single_digit')).read().strip()
    assert summarizer.summary(summary_data, 0.459329, 0.1, digits=1) == exp
    with pytest.raises(ValueError):
        summarizer.summary(summary_data, 0.1, output='test')
    assert 'p = ' in summarizer.summary(summary_data, 0.03, output='report')


REPORT_CASES = [
    # golden, rel_effect, lower, upper, alpha, p_value, digits   (reference tests/test_summary.py:78-151)
    ('test_report_summary_1', 0.41, -0.30, 0.30, 0.1, 0.5, 2),           # positive, not significant
    ('test_report_summary_2', 0.41, 0.434, 0.234, 0.1, 0.05, 2),         # positive, significant
    ('test_report_summary_3', -0.343, -0.434, 0.234, 0.1, 0.5, 2),       # negative, not significant
    ('test_report_summary_4', -0.343, -0.434, -0.234, 0.1, 0.05, 2),     # negative, significant
    ('test_report_summary_single_digit', 0.41, -0.30, 0.30, 0.1, 0.5, 1),
]


@pytest.mark.parametrize('golden,rel,lo,hi,alpha,p,digits', REPORT_CASES)
def test_report_matches_reference_goldens(summary_data, golden, rel, lo, hi, alpha, p, digits):
    """`summary(output='report')` byte for byte against the reference's five report goldens
    (/root/reference/tests/fixtures/test_report_summary_*, templates/report:1-81)."""
    summary_data = summary_data.copy()
    summary_data.loc['rel_effect', 'average'] = rel
    summary_data.loc['rel_effect_lower', 'average'] = lo
    summary_data.loc['rel_effect_upper', 'average'] = hi
    exp = open(os.path.join(GOLD, golden)).read().strip()
    assert summarizer.summary(summary_data, p, alpha=alpha, output='report', digits=digits) == exp


def test_default_model_spec(rand_data):
    data = cidata.format_input_data(rand_data)
    pre, post = cidata.process_pre_post_data(data, [0, 99], [100, 199])
    normed_pre, normed_post, mu_sig = cidata.standardize_pre_and_post_data(pre, post)
    obs = cidata._build_observed_time_series(normed_pre)
    m = cimodel.build_default_model(obs, normed_pre, normed_post, 0.01, 7, 2)
    assert [type(c).__name__ for c in m.components] == ['LocalLevel', 'SparseLinearRegression', 'Seasonal']
    assert m.param_names() == ['observation_noise_scale', 'LocalLevel/_level_scale',
                               'SparseLinearRegression/_global_scale_variance',
                               'SparseLinearRegression/_global_scale_noncentered',
                               'SparseLinearRegression/_local_scale_variances',
                               'SparseLinearRegression/_local_scales_noncentered',
                               'SparseLinearRegression/_weights_noncentered', 'Seasonal/_drift_scale']
    assert m.parameters[0].prior.family == 'SqrtInverseGamma'
    np.testing.assert_allclose(m.parameters[0].prior.params['scale'], 16 * (obs.std(ddof=0).values[0] / 2) ** 2,
                               rtol=1e-6)
    np.testing.assert_allclose(m.parameters[1].prior.params['scale'], 16 * 0.01 ** 2, rtol=1e-6)
    np.testing.assert_allclose(m.prior_level_sd, 0.01, rtol=1e-6)
    X = m.components[1].design_matrix.to_dense()
    assert X.shape == (200, 2) and X.dtype == np.float32
    np.testing.assert_array_equal(X, pd.concat([normed_pre, normed_post]).iloc[:, 1:].values.astype(np.float32))
    assert m.components[2].num_seasons == 7 and m.components[2].num_steps_per_season == 2
    cimodel.check_input_model(m, pre, post)
    with pytest.raises(ValueError) as e:
        cimodel.check_input_model('not a model', pre, post)
    assert str(e.value) == 'Input model must be of type StructuralTimeSeries.'
    with pytest.raises(ValueError) as e:
        cimodel.Sum([cimodel.LocalLevel(), object()], observed_time_series=obs)
    assert 'Unsupported structural component' in str(e.value)
    bad = cimodel.Sum([cimodel.LocalLevel(), cimodel.SparseLinearRegression(np.zeros((10, 2), np.float32))],
                      observed_time_series=obs)
    with pytest.raises(ValueError) as e:
        cimodel.check_input_model(bad, pre, post)
    assert 'Customized Linear Regression Models must have total' in str(e.value)


def test_custom_model_lowering_without_gpu():
    """Sum -> layout flags / parameter names for the wider model set (reference main.py:180-252, notebook
    cell 116): LocalLinearTrend, dense LinearRegression, up to two Seasonal components; everything else
    is rejected."""
    from tfcausalimpact_b200 import _native as nt
    obs = pd.Series(np.arange(40, dtype=np.float32))
    X = np.zeros((50, 3), np.float32)
    m = cimodel.Sum([cimodel.LocalLinearTrend(observed_time_series=obs), cimodel.LinearRegression(design_matrix=X),
                     cimodel.Seasonal(num_seasons=4, name='Month'), cimodel.Seasonal(num_seasons=52, name='Year')],
                    observed_time_series=obs)
    assert m.flags == nt.FLAG_LLT | nt.FLAG_DENSE_REG | nt.FLAG_OBS_LOGNORMAL
    assert m.param_names() == ['observation_noise_scale', 'LocalLinearTrend/_level_scale',
                               'LocalLinearTrend/_slope_scale', 'LinearRegression/_weights',
                               'Month/_drift_scale', 'Year/_drift_scale']
    assert nt.num_params(3, 4, m.flags, 52) == 1 + 2 + 3 + 2
    assert [(n, b - a) for n, a, b, _ in m.param_slices()][3] == ('LinearRegression/_weights', 3)
    with pytest.raises(ValueError):      # three Seasonal components
        cimodel.Sum([cimodel.LocalLevel(), cimodel.Seasonal(2, name='a'), cimodel.Seasonal(3, name='b'),
                     cimodel.Seasonal(4, name='c')], observed_time_series=obs)
    with pytest.raises(ValueError):      # same name twice
        cimodel.Sum([cimodel.LocalLevel(), cimodel.Seasonal(4), cimodel.Seasonal(7)], observed_time_series=obs)
    with pytest.raises(ValueError):      # seasonal before the trend
        cimodel.Sum([cimodel.Seasonal(4), cimodel.LocalLevel()], observed_time_series=obs)
    with pytest.raises(ValueError):      # two trends
        cimodel.Sum([cimodel.LocalLevel(), cimodel.LocalLinearTrend()], observed_time_series=obs)


def test_no_cpu_fallback(monkeypatch):
    """The product path must fail loudly without CUDA."""
    import torch
    from tfcausalimpact_b200 import _native
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    with pytest.raises(_native.NativeError):
        _native.require_cuda()
    from tfcausalimpact_b200.engine import DeviceBatch
    with pytest.raises(_native.NativeError):
        DeviceBatch(np.zeros((1, 10), np.float32))


@pytest.mark.parametrize('freq,hole,expect_len', [
    ('MS', 7, 24),        # month starts, one missing month -> exactly one NaN row (was re-gridded at 28 days)
    ('QS', 3, 12),        # quarter starts
    ('YS', 2, 8),         # year starts
    ('ME', 5, 24),        # month ends
    ('D', 10, 40),        # daily
    ('h', 10, 40),        # hourly
    ('W-SUN', 4, 20),     # weekly
])
def test_regularize_series_calendar_steps_with_a_hole(freq, hole, expect_len):
    idx = pd.date_range('2015-01-01', periods=expect_len, freq=freq)
    df = pd.DataFrame({'y': np.arange(expect_len, dtype=float)}, index=idx).drop(idx[hole])
    assert df.index.freq is None
    out = cidata.regularize_series(df)
    assert len(out) == expect_len and out.index.equals(idx)
    assert int(out['y'].isna().sum()) == 1 and np.isnan(out['y'].iloc[hole])
    assert out['y'].notna().sum() == df['y'].notna().sum()


def test_regularize_series_business_days_and_irregular_stamps():
    bd = pd.bdate_range('2020-01-06', periods=10)                     # Mon-Fri x 2: smallest gap = 1 day
    out = cidata.regularize_series(pd.DataFrame({'y': np.arange(10.0)}, index=pd.DatetimeIndex(bd.values)))
    assert len(out) == 12 and int(out['y'].isna().sum()) == 2          # the week-end becomes two masked steps
    irr = pd.DatetimeIndex(['2020-01-01 00:00', '2020-01-01 00:07', '2020-01-01 00:17', '2020-01-01 01:00'])
    df = pd.DataFrame({'y': [1.0, 2.0, 3.0, 4.0]}, index=irr)
    with pytest.warns(RuntimeWarning):
        out = cidata.regularize_series(df)
    assert out is df                                                   # never drops observations silently


def _batch_frames(N=5, T=40, gap=(7, 33)):
    rng = np.random.default_rng(2)
    full = pd.date_range('2020-01-01', periods=T, freq='D')
    keep = np.ones(T, dtype=bool)
    keep[list(gap)] = False
    idx = pd.DatetimeIndex(full[keep].values)
    return [pd.DataFrame(rng.normal(size=(keep.sum(), 3)), index=idx, columns=['y', 'a', 'b']) for _ in range(N)], full


def test_process_batch_input_groups_gaps_and_periods():
    """Vectorised front door: one group per shared index, date periods resolved once, gaps -> NaN rows
    of every series; equal to the single-series `process_input_data` slices."""
    from tfcausalimpact_b200.batch_input import process_batch_input
    frames, full = _batch_frames()
    odd = frames[0].copy()
    odd.index = pd.DatetimeIndex(odd.index.values) + pd.Timedelta(days=0)      # equal but distinct index object
    other = pd.DataFrame(np.random.default_rng(3).normal(size=(40, 3)), index=full, columns=['y', 'a', 'b'])
    proc = process_batch_input(frames + [odd, other], ['2020-01-01', '2020-01-30'], ['2020-01-31', '2020-02-09'])
    assert proc['n'] == 7 and len(proc['groups']) == 2
    g = proc['groups'][0]
    assert list(g.positions) == [0, 1, 2, 3, 4, 5]
    y_pre, y_post, X = (t.numpy() for t in g.tensors('cpu'))            # the same torch code runs on the device
    assert y_pre.shape == (6, 30) and y_post.shape == (6, 10) and X.shape == (6, 40, 2)
    assert np.isnan(y_pre[:, 7]).all() and np.isnan(y_post[:, 3]).all() and np.isnan(X[:, 7]).all()
    ref = cidata.process_input_data(frames[2], ['2020-01-01', '2020-01-30'], ['2020-01-31', '2020-02-09'], None,
                                    {'standardize': False}, 0.05)
    np.testing.assert_allclose(y_pre[2], ref['pre_data'].iloc[:, 0].values, rtol=1e-6, equal_nan=True)
    np.testing.assert_allclose(y_post[2], ref['post_data'].iloc[:, 0].values, rtol=1e-6, equal_nan=True)
    np.testing.assert_allclose(X[2, :30], ref['pre_data'].iloc[:, 1:].values, rtol=1e-6, equal_nan=True)
    assert g.pre_index.equals(ref['pre_data'].index) and g.post_index.equals(ref['post_data'].index)
    y1 = proc['groups'][1].tensors('cpu')[0].numpy()
    assert y1.shape == (1, 30) and not np.isnan(y1).any()


def test_process_batch_input_errors_name_the_series():
    from tfcausalimpact_b200.batch_input import BatchInputError
    from tfcausalimpact_b200 import batch_input

    def process_batch_input(*a):                    # periods / stacking on the host, per-series checks where the block lives
        proc = batch_input.process_batch_input(*a)
        for g in proc['groups']:
            g.tensors('cpu')
        return proc
    frames, _ = _batch_frames()
    pre, post = ['2020-01-01', '2020-01-30'], ['2020-01-31', '2020-02-09']
    bad = [f.copy() for f in frames]
    bad[3].iloc[:, 0] = 1.0
    with pytest.raises(BatchInputError) as e:
        process_batch_input(bad, pre, post)
    assert str(e.value) == 'Input response cannot be constant.' and e.value.series == 3
    bad = [f.copy() for f in frames]
    bad[1].iloc[4, 2] = np.nan
    with pytest.raises(ValueError) as e:
        process_batch_input(bad, pre, post)
    assert str(e.value) == 'Input data cannot have NAN values.' and e.value.series == 1
    bad = [f.copy() for f in frames]
    bad[2].iloc[:, 0] = np.nan
    with pytest.raises(ValueError) as e:
        process_batch_input(bad, pre, post)
    assert str(e.value) == 'Input response cannot have just Null values.'
    bad = [f.copy().astype(object) for f in frames[:2]]
    bad[1].iloc[0, 1] = 'x'
    with pytest.raises(ValueError) as e:
        process_batch_input(bad, pre, post)
    assert str(e.value) == 'Input data must contain only numeric values.' and e.value.series == 1
    with pytest.raises(ValueError) as e:
        process_batch_input(frames, pre, ['2020-01-30', '2020-02-09'])
    assert 'Values in training data' in str(e.value) or 'post_period first value' in str(e.value)
    with pytest.raises(ValueError) as e:
        process_batch_input(frames, pre, ['2020-03-30', '2020-04-09'])
    assert str(e.value) == '2020-03-30 not present in input data index.'
    arr = np.random.default_rng(0).normal(size=(4, 30, 2)).astype(np.float32)
    proc = process_batch_input(arr, [0, 19], [20, 29])
    assert proc['groups'][0].tensors('cpu')[0].shape == (4, 20) and proc['groups'][0].range_index
    with pytest.raises(ValueError) as e:
        process_batch_input(arr, [0, 1], [20, 29])
    assert str(e.value) == 'pre_period must span at least 3 time points.'
