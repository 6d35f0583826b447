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
    cimodel.check_capability_limits(model_args)
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


def _infer_step(idx: pd.DatetimeIndex):
    """Grid step of a DatetimeIndex the way `tfp.sts.regularize_series` infers it (documented behaviour: the
    smallest distance between consecutive observations, with months and years treated as calendar units):
    the index's own `freq` if it carries one; else, when every stamp falls on the same day-of-month (or every
    stamp is a month end) at the same time of day, the smallest month distance as a calendar offset; else the
    smallest positive gap as a fixed-length step.  `pd.infer_freq` is deliberately NOT used: it returns
    business-day / anchored frequencies ('B', 'W-SUN', ...) that would hide gaps the smallest-gap rule fills
    with NaN (a Mon-Fri index gets NaN week-ends, as with TFP)."""
    if idx.freq is not None:
        return idx.freq
    gaps = np.diff(idx.values)
    pos = gaps[gaps > np.timedelta64(0)]
    if len(pos) == 0:
        return None
    same_clock = bool((idx.normalize() + (idx[0] - idx[0].normalize()) == idx).all())
    months = np.diff(idx.year.values * 12 + idx.month.values)
    if same_clock and (months > 0).any() and pd.Timedelta(pos.min()) >= pd.Timedelta(days=28):
        m = int(months[months > 0].min())
        if bool((months % m == 0).all()):
            if bool((idx.day == idx[0].day).all()):
                return pd.DateOffset(years=m // 12) if m % 12 == 0 else pd.DateOffset(months=m)
            if bool(idx.is_month_end.all()):
                return pd.offsets.MonthEnd(m)
    step = pd.Timedelta(pos.min())
    return pd.DateOffset(days=step.days) if step == pd.Timedelta(days=step.days) and step.days > 0 else step


def regularize_series(frame: pd.DataFrame) -> pd.DataFrame:
    """Re-index on a regular grid, NaN where the input had gaps (stand-in for
    `tfp.sts.regularize_series`, /root/reference/causalimpact/data.py:333-334).  If the inferred grid would
    DROP observations (stamps that are not multiples of the step) the frame is returned unchanged with a
    warning instead of silently fitting a mostly-NaN series."""
    idx = frame.index
    if not isinstance(idx, pd.DatetimeIndex) or len(idx) < 2:
        return frame
    step = _infer_step(idx)
    if step is None:
        return frame
    out = frame.asfreq(step)
    if int(out.iloc[:, 0].notna().sum()) < int(frame.iloc[:, 0].notna().sum()):
        import warnings
        warnings.warn('regularize_series: the time index is not a regular grid with gaps (step '
                      f'{step} would drop observations); the series is used as given.', RuntimeWarning)
        return frame
    return out


def process_pre_post_data(data: pd.DataFrame, pre_period, post_period) -> List[pd.DataFrame]:
    pre = process_period(pre_period, data)
    post = process_period(post_period, data)
    check_periods(pre, post, pre_period, post_period)
    if isinstance(data.index, pd.RangeIndex):
        # the state-space model needs a calendar: synthesise a daily one
        data = data.set_index(pd.date_range(start='2020-01-01', periods=len(data)))
    pre_data = data.iloc[pre[0]: pre[1] + 1, :]
    post_data = data.iloc[post[0]: post[1] + 1, :]
    return [regularize_series(pre_data), regularize_series(post_data)]


def check_periods(pre: List[int], post: List[int], pre_period, post_period) -> None:
    """Consistency rules of the two periods as integer positions (data.py:360-385 of the reference)."""
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
