"""Input validation and slicing for `CausalImpact(...)`.

Host-side pandas work only (O(T) per series); behaviour -- accepted types, slicing rules and the
exact error strings -- follows /root/reference/causalimpact/data.py:31-496 so the reference's own
tests/test_data.py expectations hold.  The one TFP call there (`tfp.sts.regularize_series`,
data.py:333-334) is replaced by a pandas `asfreq` re-indexing (the reference's tests pin the two as
equivalent: tests/test_data.py:52-58).
"""
from typing import Any, Dict, List, Optional, Tuple, Union

import numpy as np
import pandas as pd

from . import model as cimodel
from .misc import standardize

_EMPTY_EXEMPT = {'model'}


def process_input_data(data, pre_period, post_period, model, model_args, alpha) -> Dict[str, Any]:
    """Validate everything the constructor received and return the 13-key dict the orchestrator
    consumes (data.py:151-165 of the reference)."""
    _check_empty_inputs({'data': data, 'pre_period': pre_period, 'post_period': post_period,
                         'model': model, 'model_args': model_args, 'alpha': alpha})
    data = format_input_data(data)
    pre_data, post_data = process_pre_post_data(data, pre_period, post_period)
    mask = _build_nan_mask(post_data)
    alpha = process_alpha(alpha)
    model_args = cimodel.process_model_args(model_args if model_args else {})
    if model_args['standardize']:
        normed_pre, normed_post, mu_sig = standardize_pre_and_post_data(pre_data, post_data)
    else:
        normed_pre = normed_post = mu_sig = None
    observed_time_series = _build_observed_time_series(pre_data if normed_pre is None else normed_pre)
    if model:
        cimodel.check_input_model(model, pre_data, post_data)
    else:
        model = cimodel.build_default_model(
            observed_time_series,
            normed_pre if model_args['standardize'] else pre_data,
            normed_post if model_args['standardize'] else post_data,
            model_args['prior_level_sd'], model_args['nseasons'], model_args['season_duration'])
    return {
        'data': data, 'pre_period': pre_period, 'post_period': post_period,
        'pre_data': pre_data, 'post_data': post_data,
        'normed_pre_data': normed_pre, 'normed_post_data': normed_post,
        'observed_time_series': observed_time_series, 'model': model, 'model_args': model_args,
        'alpha': alpha, 'mu_sig': mu_sig, 'mask': mask,
    }


def _build_nan_mask(post_data: pd.DataFrame) -> np.ndarray:
    """True where the post-period response is present (regularisation can insert NaN rows)."""
    return post_data.iloc[:, 0].notna().values


def _build_observed_time_series(pre_data: pd.DataFrame) -> pd.DataFrame:
    return pd.DataFrame(pre_data.iloc[:, 0])


def _check_empty_inputs(inputs: Dict[str, Any]) -> None:
    missing = sorted(k for k, v in inputs.items() if v is None and k not in _EMPTY_EXEMPT)
    if missing:
        plural = 's' if len(missing) > 1 else ''
        raise ValueError(f'{", ".join(missing)} input argument{plural} cannot be empty')


def standardize_pre_and_post_data(pre_data: pd.DataFrame, post_data: pd.DataFrame
                                  ) -> Tuple[pd.DataFrame, pd.DataFrame, Tuple[float, float]]:
    """Standardise both periods with the PRE-period mean/std of every column; mu_sig is the
    response's (mean, std)."""
    normed_pre, (mu, sig) = standardize(pre_data)
    normed_post = (post_data - mu) / sig
    return normed_pre, normed_post, (mu.iloc[0], sig.iloc[0])


def format_input_data(data) -> pd.DataFrame:
    if not isinstance(data, pd.DataFrame):
        try:
            data = pd.DataFrame(data)
        except ValueError:
            raise ValueError('Could not transform input data to pandas DataFrame.')
    validate_y(data.iloc[:, 0])
    elementwise = data.map if hasattr(data, 'map') else data.applymap
    if not elementwise(np.isreal).values.all():
        raise ValueError('Input data must contain only numeric values.')
    if data.shape[1] > 1 and data.iloc[:, 1:].isna().values.any():
        raise ValueError('Input data cannot have NAN values.')
    data = convert_index_to_datetime(data)
    return data.astype(np.float32)


def regularize_series(frame: pd.DataFrame) -> pd.DataFrame:
    """Re-index on a regular grid, NaN where the input had gaps.

    Stand-in for `tfp.sts.regularize_series`: use the index's own frequency, else pandas'
    inferred one, else the smallest positive gap between consecutive stamps."""
    idx = frame.index
    if not isinstance(idx, pd.DatetimeIndex) or len(idx) < 2:
        return frame
    freq = idx.freq
    if freq is None:
        try:
            freq = pd.infer_freq(idx) if len(idx) >= 3 else None
        except (TypeError, ValueError):
            freq = None
    if freq is None:
        gaps = np.diff(idx.values)
        gaps = gaps[gaps > np.timedelta64(0)]
        if len(gaps) == 0:
            return frame
        step = pd.Timedelta(gaps.min())
        freq = pd.DateOffset(days=step.days) if step == pd.Timedelta(days=step.days) and step.days > 0 else step
    return frame.asfreq(freq)


def process_pre_post_data(data: pd.DataFrame, pre_period, post_period) -> List[pd.DataFrame]:
    pre = process_period(pre_period, data)
    post = process_period(post_period, data)
    if pre[1] > post[0]:
        raise ValueError(
            'Values in training data cannot be present in the post-intervention '
            'data. Please fix your pre_period value to cover at most one point less '
            'from when the intervention happened.')
    if pre[1] < pre[0]:
        raise ValueError('pre_period last number must be bigger than its first.')
    if pre[1] - pre[0] < 3:
        raise ValueError('pre_period must span at least 3 time points.')
    if post[1] < post[0]:
        raise ValueError('post_period last number must be bigger than its first.')
    if post[0] <= pre[1]:
        raise ValueError(f'post_period first value ({post_period[0]}) must '
                         'be bigger than the second value of pre_period '
                         f'({pre_period[1]}).')
    if isinstance(data.index, pd.RangeIndex):
        # the state-space model needs a calendar: synthesise a daily one
        data = data.set_index(pd.date_range(start='2020-01-01', periods=len(data)))
    pre_data = data.iloc[pre[0]: pre[1] + 1, :]
    post_data = data.iloc[post[0]: post[1] + 1, :]
    return [regularize_series(pre_data), regularize_series(post_data)]


def validate_y(y: pd.Series) -> None:
    if np.all(y.isna()):
        raise ValueError('Input response cannot have just Null values.')
    if y.notna().values.sum() < 3:
        raise ValueError('Input response must have more than 3 non-null points at least.')
    if y.std(skipna=True, ddof=0) == 0:
        raise ValueError('Input response cannot be constant.')


def convert_index_to_datetime(data: pd.DataFrame) -> pd.DataFrame:
    """String date labels ('20200101', ...) become a DatetimeIndex when they parse."""
    if isinstance(data.index.values[0], str):
        try:
            data.set_index(pd.to_datetime(data.index), inplace=True)
        except ValueError:
            pass
    return data


def process_period(period, data: pd.DataFrame) -> List[int]:
    if not isinstance(period, list):
        raise ValueError('Input period must be of type list.')
    if len(period) != 2:
        raise ValueError('Period must have two values regarding the beginning and end of '
                         'the pre and post intervention data.')
    if any(p is None for p in period):
        raise ValueError('Input period cannot have `None` values.')
    # NB: like the reference (data.py:423-427) only period[1] is inspected for str / Timestamp
    both_int = isinstance(period[0], int) and isinstance(period[1], int)
    if not (both_int or isinstance(period[1], str) or isinstance(period[1], pd.Timestamp)):
        raise ValueError('Input periods must contain either int, str or pandas Timestamp')
    for point in period:
        if point not in data.index:
            label = point.strftime('%Y%m%d') if isinstance(point, pd.Timestamp) else str(point)
            raise ValueError(f'{label} not present in input data index.')
    if isinstance(period[0], (str, pd.Timestamp)):
        period = convert_date_period_to_int(period, data)
    return period


def convert_date_period_to_int(period, data: pd.DataFrame) -> List[int]:
    return [data.index.get_loc(date) for date in period]


def process_alpha(alpha: float) -> float:
    if not isinstance(alpha, float):
        raise ValueError('alpha must be of type float.')
    if alpha < 0 or alpha > 1:
        raise ValueError('alpha must range between 0 (zero) and 1 (one) inclusive.')
    return alpha
