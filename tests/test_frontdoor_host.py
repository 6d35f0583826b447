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
