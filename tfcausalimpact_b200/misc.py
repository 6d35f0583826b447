"""Small numeric helpers of the front door (mirrors /root/reference/causalimpact/misc.py:27-126)."""
from typing import Optional, Tuple

import pandas as pd
from scipy.special import ndtri


def standardize(data: pd.DataFrame):
    """Column-wise (x - mean) / std with population std (ddof=0); NaN std -> 1.

    Same contract as the reference (misc.py:27-54): returns [normed, (mu, std)] and raises when
    the frame has a single row."""
    if data.shape[0] == 1:
        raise ValueError('Input data must have more than one value')
    mu = data.mean(skipna=True)
    std = data.std(skipna=True, ddof=0)
    return [(data - mu) / std.fillna(1), (mu, std)]


def unstandardize(data, mus_sigs: Tuple[float, float]):
    mu, sig = mus_sigs
    return (data * sig) + mu


def maybe_unstandardize(data, mu_sig: Optional[Tuple[float, float]] = None):
    """Identity when no standardisation was applied (mu_sig is None); misc.py:80-107."""
    return data if mu_sig is None else unstandardize(data, mu_sig)


def get_z_score(p: float) -> float:
    """Standard-normal quantile (the reference asks tfd.Normal(0, 1).quantile, misc.py:110-126)."""
    return float(ndtri(p))
