"""Sample reductions of the front door, computed on the GPU (K8 kernels through the C-ABI).

Same public functions and result frames as /root/reference/causalimpact/inferences.py
(compile_posterior_inferences :52, summarize_posterior_inferences :285, compute_p_value :372,
build_cum_index :252, get_lower_upper_percentiles :34); the NumPy percentile / cumsum / counting
work is done by `ci_reduce_inferences` / `ci_reduce_summary`.  The distribution arguments are duck
typed exactly as in the reference (`.sample(n)` -> array-like with `.numpy()`, `.mean()`); native
distributions additionally expose `sample_device` so samples never visit the host.
"""
from typing import List, Optional, Tuple

import numpy as np
import pandas as pd
import torch

from . import engine
from . import _native as nt

COLUMNS = [
    'complete_preds_means', 'complete_preds_lower', 'complete_preds_upper',
    'post_preds_means', 'post_preds_lower', 'post_preds_upper',
    'post_cum_y', 'post_cum_preds_means', 'post_cum_preds_lower', 'post_cum_preds_upper',
    'point_effects_means', 'point_effects_lower', 'point_effects_upper',
    'post_cum_effects_means', 'post_cum_effects_lower', 'post_cum_effects_upper',
]
_COMPLETE = {0, 1, 2, 10, 11, 12}
_POST = {3, 4, 5}
SUMMARY_INDEX = ['actual', 'predicted', 'predicted_lower', 'predicted_upper', 'abs_effect',
                 'abs_effect_lower', 'abs_effect_upper', 'rel_effect', 'rel_effect_lower', 'rel_effect_upper']


def get_lower_upper_percentiles(alpha: float) -> List[float]:
    return [alpha * 100. / 2., 100 - alpha * 100. / 2.]


def _device():
    nt.require_cuda()
    return torch.device(f'cuda:{torch.cuda.current_device()}')


def _to_device(a, dev) -> torch.Tensor:
    if torch.is_tensor(a):
        return a.to(dev, torch.float32).contiguous()
    if hasattr(a, 'numpy') and not isinstance(a, np.ndarray):
        a = a.numpy()
    return torch.as_tensor(np.array(a, dtype=np.float32, order="C")).to(dev)


def _samples(dist, niter, mu_sig, dev, width):
    """[niter, width] float32 on the device, original units."""
    if hasattr(dist, 'sample_device'):
        ms = None if mu_sig is None else torch.tensor([[mu_sig[0]], [mu_sig[1]]], dtype=torch.float32, device=dev)
        return dist.sample_device(niter, ms), dist.mean_device(ms)
    s = _to_device(dist.sample(niter), dev).reshape(niter, width)
    m = _to_device(dist.mean(), dev).reshape(width)
    if mu_sig is not None:
        s = s * float(mu_sig[1]) + float(mu_sig[0])
        m = m * float(mu_sig[1]) + float(mu_sig[0])
    return s.contiguous(), m.contiguous()


def compile_posterior_inferences(original_index, mask, pre_data: pd.DataFrame, post_data: pd.DataFrame,
                                 one_step_dist, posterior_dist, mu_sig: Optional[Tuple[float, float]],
                                 alpha: float = 0.05, niter: int = 1000) -> pd.DataFrame:
    """The 16-column inference frame (columns / order of inferences.py:229-246)."""
    dev = _device()
    T, Tf = len(pre_data), len(post_data)
    mask = np.asarray(mask, dtype=bool)
    y_pre = _to_device(pre_data.iloc[:, 0].values, dev)[None, :]
    y_post_np = post_data.iloc[:, 0].values.astype(np.float32).copy()
    y_post_np[~mask] = np.nan
    y_post = _to_device(y_post_np, dev)[None, :]
    pre_sims, pre_mean = _samples(one_step_dist, niter, mu_sig, dev, T)
    post_sims, post_mean = _samples(posterior_dist, niter, mu_sig, dev, Tf)
    out = engine.reduce_inferences(y_pre, y_post, pre_sims[None], pre_mean[None], post_sims[None],
                                   post_mean[None], float(alpha))[0].cpu().numpy().astype(np.float64)
    return frame_from_reduction(out, mask, pre_data.index, post_data.index, isinstance(original_index, pd.RangeIndex))


def frame_from_reduction(out: np.ndarray, mask: np.ndarray, pre_index: pd.Index, post_index: pd.Index,
                         range_index: bool) -> pd.DataFrame:
    """[16, T + Tf + 1] device reduction (masked post steps compacted to the front) -> the reference's 16-column
    frame: complete / point-effect columns over pre + observed post labels, post columns over observed post labels,
    cumulative columns prefixed by a 0 at the last pre label (inferences.py:208-282)."""
    T = len(pre_index)
    post_masked = post_index[np.asarray(mask, dtype=bool)]
    nvalid = len(post_masked)
    full_index = pre_index.append(post_masked)
    cum_index = build_cum_index(pre_index, post_masked)
    frame = pd.DataFrame(np.nan, index=full_index, columns=COLUMNS, dtype=np.float64)
    for c, name in enumerate(COLUMNS):
        if c in _COMPLETE:
            frame[name] = out[c, :T + nvalid]
        elif c in _POST:
            frame.loc[post_masked, name] = out[c, :nvalid]
        else:
            frame.loc[cum_index, name] = out[c, :nvalid + 1]
    if range_index:
        frame.set_index(pd.RangeIndex(start=0, stop=len(frame)), inplace=True)
    return frame


def build_cum_index(pre_data_index, post_data_index):
    """Post index extended with the last pre-period label (the zero-prefixed cumulative columns
    start there; inferences.py:252-282)."""
    dtype = post_data_index.dtype
    return post_data_index.union([pre_data_index[-1]]).astype(dtype)


def _summary_device(y_post, post_mean, sims, alpha):
    dev = _device()
    y = _to_device(y_post, dev).reshape(1, -1)
    m = _to_device(post_mean, dev).reshape(1, -1)
    s = _to_device(sims, dev)
    s = s.reshape(1, s.shape[0], -1)
    return engine.reduce_summary(y, m, s, float(alpha))[0].cpu().numpy().astype(np.float64)


def summarize_posterior_inferences(post_preds_means, post_data: pd.DataFrame, simulated_ys,
                                   alpha: float = 0.05) -> pd.DataFrame:
    """10x2 table (average, cumulative) of actual / predicted / effects (inferences.py:311-369).
    `post_preds_means`, `post_data` and `simulated_ys` are already restricted to observed steps."""
    means = np.asarray(post_preds_means, dtype=np.float32)
    means = means[~np.isnan(means)]
    out = _summary_device(post_data.iloc[:, 0].values, means, simulated_ys, alpha)
    return pd.DataFrame(out[:20].reshape(10, 2), columns=['average', 'cumulative'], index=SUMMARY_INDEX)


def compute_p_value(simulated_ys, post_data_sum: float) -> float:
    """min(#{row sum > total}, #{row sum < total}) / (n + 1), strict (inferences.py:397-408)."""
    s = simulated_ys
    width = s.shape[1] if len(s.shape) > 1 else 1
    y = np.zeros(width, dtype=np.float32)
    y[0] = post_data_sum
    out = _summary_device(y, np.ones(width, dtype=np.float32), s, 0.05)
    return float(out[20])
