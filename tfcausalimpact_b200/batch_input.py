"""Vectorised input processing for the batched front door.

The reference validates, slices and regularises ONE DataFrame per `CausalImpact(...)` call
(/root/reference/causalimpact/data.py:31-165: `format_input_data`, `process_pre_post_data`,
`tfp.sts.regularize_series`, `standardize_pre_and_post_data`).  At 10k series that per-series pandas work is the
wall-clock bottleneck once the fit runs on the GPU, so here it is done once per GROUP of series that share a time
index:

  * frames are stacked into one float32 block [N, T_raw, 1 + k] (one `to_numpy` per frame, no per-series pandas
    arithmetic); a long-format frame sorted by (series, time) is reshaped without any copy per series
  * periods (int positions, date strings or Timestamps) are resolved against the group's index with the reference's
    `process_period` / `process_pre_post_data` rules and error strings, once per group
  * the pre and post slices are regularised once per group (`data.regularize_series` on the index); the resulting
    position map scatters all series of the group into NaN-initialised grids, so a gap becomes a masked step of
    every series -- the NaN mask is data, built with one vectorised assignment
  * the block is uploaded ONCE; per-series validation (`validate_y`, NaN-covariate check), the scatter into the grids and
    the pre/post concatenation of the covariates run on the device (`SeriesGroup.tensors`); a failed check raises
    `BatchInputError` (a ValueError carrying the reference's exact message and the offending series position)
  * standardisation happens on the device (`batch.standardize_batch`).
"""
from typing import Any, Dict, List, Optional, Sequence

import numpy as np
import pandas as pd

from . import data as cidata


class BatchInputError(ValueError):
    """ValueError with the reference's message for the single-series check that failed; `.series` is the position
    of the first offending series in the batch."""

    def __init__(self, message: str, series: int):
        super().__init__(message)
        self.series = int(series)


class SeriesGroup:
    """Series that share a time index: the raw float32 block [n, T_raw, 1 + k] plus everything needed to slice and
    regularise it -- period positions and the maps of raw rows into the regular pre / post grids.  `tensors(device)`
    materialises y_pre [n, T], y_post [n, Tf] (NaN = masked step) and X [n, T + Tf, k] (NaN rows at inserted steps)
    ON THE DEVICE: the block is uploaded once and validated, scattered and concatenated there, so the host touches
    every value exactly once (when the frames are stacked)."""

    def __init__(self, positions, block, pre, post, pre_pos, post_pos, pre_index, post_index, range_index):
        self.positions = np.asarray(positions)      # positions of these series in the caller's batch
        self.block = block
        self.pre, self.post = pre, post
        self.pre_pos, self.post_pos = pre_pos, post_pos
        self.pre_index, self.post_index = pre_index, post_index
        self.range_index = range_index      # the input had a RangeIndex (inferences are re-indexed 0..n-1)
        self.T, self.Tf, self.k = len(pre_index), len(post_index), block.shape[2] - 1

    @property
    def shape_key(self):
        return (self.T, self.Tf, self.k)

    def tensors(self, device, lo: int = 0, hi: Optional[int] = None):
        """(y_pre, y_post, X) of series lo..hi of the group as float32 tensors on `device`; validates them with the
        reference's per-series rules (`validate_y`, covariate NaN check: data.py:268-270,280-294)."""
        import torch
        dev = torch.device(device)
        blk = torch.from_numpy(self.block[lo:hi])
        if dev.type == 'cuda':
            blk = blk.to(dev, non_blocking=True)         # asynchronous when the block is page-locked (_alloc_block)
        _validate_block(blk, self.positions[lo:hi])

        def grid(rows, pos, length):
            if pos is None:
                return rows
            out = torch.full((rows.shape[0], length, rows.shape[2]), float('nan'), dtype=torch.float32, device=dev)
            out[:, torch.as_tensor(pos, device=dev)] = rows
            return out
        pre = grid(blk[:, self.pre[0]: self.pre[1] + 1], self.pre_pos, self.T)
        post = grid(blk[:, self.post[0]: self.post[1] + 1], self.post_pos, self.Tf)
        X = torch.cat([pre[:, :, 1:], post[:, :, 1:]], dim=1).contiguous() if self.k > 0 else None
        return pre[:, :, 0].contiguous(), post[:, :, 0].contiguous(), X


def _alloc_block(shape) -> np.ndarray:
    """float32 staging block; page-locked when a CUDA device is present, so the single upload of the block is an
    asynchronous DMA straight from it (the NumPy array is a view of the pinned torch tensor)."""
    import torch
    if torch.cuda.is_available():
        try:
            return torch.empty(shape, dtype=torch.float32, pin_memory=True).numpy()
        except RuntimeError:
            pass
    return np.empty(shape, dtype=np.float32)


def _to_block(frames: Sequence[pd.DataFrame], positions: Sequence[int]) -> np.ndarray:
    """Stack the frames' values into one [n, T_raw, 1 + k] float32 block.  pandas keeps a frame column-major, so every
    copy is a small transpose; the copies release the GIL and are spread over a few host threads."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    n, (t, c) = len(frames), frames[0].shape
    block = _alloc_block((n, t, c))

    def work(lo, hi):
        for i in range(lo, hi):
            f = frames[i]
            if f.shape != (t, c):
                raise BatchInputError(f'series {positions[i]} has shape {f.shape}, expected {(t, c)} '
                                      '(series sharing an index must share the columns).', positions[i])
            try:
                block[i] = f.to_numpy(dtype=np.float32, copy=False)
            except (ValueError, TypeError):
                raise BatchInputError('Input data must contain only numeric values.', positions[i])
    nthr = max(1, min(16, os.cpu_count() or 1, n // 256))
    if nthr == 1:
        work(0, n)
    else:
        step = (n + nthr - 1) // nthr
        with ThreadPoolExecutor(nthr) as ex:
            futs = [ex.submit(work, lo, min(lo + step, n)) for lo in range(0, n, step)]
            errs = [f.exception() for f in futs]
        errs = [e for e in errs if e is not None]
        if errs:
            raise min(errs, key=lambda e: getattr(e, 'series', 1 << 62))
    return block


def _first(flag) -> int:
    """Position of the first True of a 1-D boolean tensor, -1 if none (one tiny device -> host read)."""
    idx = flag.nonzero()
    return int(idx[0]) if idx.numel() else -1


def _validate_block(blk, positions: Sequence[int]) -> None:
    """`validate_y` + covariate NaN check, vectorised over the series axis (torch: runs where the block lives)."""
    import torch
    y = blk[:, :, 0]
    present = ~torch.isnan(y)
    cnt = present.sum(dim=1)
    bad = _first(cnt == 0)
    if bad >= 0:
        raise BatchInputError('Input response cannot have just Null values.', positions[bad])
    bad = _first(cnt < 3)
    if bad >= 0:
        raise BatchInputError('Input response must have more than 3 non-null points at least.', positions[bad])
    inf = torch.tensor(float('inf'), device=y.device)
    lo = torch.where(present, y, inf).amin(dim=1)
    hi = torch.where(present, y, -inf).amax(dim=1)
    bad = _first(lo == hi)
    if bad >= 0:
        raise BatchInputError('Input response cannot be constant.', positions[bad])
    if blk.shape[2] > 1:
        bad = _first(torch.isnan(blk[:, :, 1:]).flatten(1).any(dim=1))
        if bad >= 0:
            raise BatchInputError('Input data cannot have NAN values.', positions[bad])


def _grid_positions(raw_index: pd.Index):
    """Regular grid of a slice's index + the position of every raw row in it (one pandas pass per GROUP)."""
    probe = pd.DataFrame({'p': np.arange(len(raw_index), dtype=np.float64)}, index=raw_index)
    grid = cidata.regularize_series(probe)
    if grid is probe or len(grid) == len(raw_index):
        return raw_index, None
    pos = np.flatnonzero(grid['p'].notna().values)
    return grid.index, pos


def _build_group(block: np.ndarray, index: pd.Index, positions, pre_period, post_period) -> SeriesGroup:
    probe = pd.DataFrame({'y': np.zeros(len(index))}, index=index)
    probe = cidata.convert_index_to_datetime(probe)
    index = probe.index
    pre = cidata.process_period(pre_period, probe)
    post = cidata.process_period(post_period, probe)
    cidata.check_periods(pre, post, pre_period, post_period)
    range_index = isinstance(index, pd.RangeIndex)
    if range_index:
        index = pd.date_range(start='2020-01-01', periods=len(index))     # data.py:326-327
    pre_index, pre_pos = _grid_positions(index[pre[0]: pre[1] + 1])
    post_index, post_pos = _grid_positions(index[post[0]: post[1] + 1])
    return SeriesGroup(positions, block, pre, post, pre_pos, post_pos, pre_index, post_index, range_index)


def _from_long(data: pd.DataFrame, series_col: str, time_col: Optional[str]):
    """Long-format frame -> list of (index, block, positions).  Rows sorted by (series, time) with the same stamps
    in every series become ONE block by a reshape."""
    keys = [series_col] + ([time_col] if time_col else [])
    codes, uniques = pd.factorize(data[series_col], sort=False)       # codes in order of first appearance
    times = data[time_col].to_numpy() if time_col else None
    dc = np.diff(codes)
    ordered = bool((dc >= 0).all()) and (times is None or bool(((times[1:] >= times[:-1]) | (dc > 0)).all()))
    value_cols = [c for c in data.columns if c not in keys]
    try:
        vals = data[value_cols].to_numpy(dtype=np.float32)
    except (ValueError, TypeError):
        raise BatchInputError('Input data must contain only numeric values.', 0)
    if not ordered:
        order = np.lexsort((times, codes)) if times is not None else np.argsort(codes, kind='stable')
        codes, vals = codes[order], vals[order]
        times = times[order] if times is not None else None
    counts = np.bincount(codes)
    n = len(uniques)
    if (counts == counts[0]).all():
        t = int(counts[0])
        same = True
        if times is not None:
            tt = times.reshape(n, t)
            same = bool((tt == tt[0]).all())
        if same:
            index = pd.Index(times[:t]) if times is not None else pd.RangeIndex(t)
            return [(index, vals.reshape(n, t, len(value_cols)), np.arange(n))], list(uniques)
    out, start = [], 0
    groups: Dict[Any, List[int]] = {}
    for i, c in enumerate(counts):
        key = (int(c), times[start:start + c].tobytes() if times is not None else b'')
        groups.setdefault(key, []).append(i)
        start += c
    offs = np.concatenate([[0], np.cumsum(counts)])
    for (c, _), members in groups.items():
        blk = np.stack([vals[offs[i]:offs[i] + c] for i in members])
        index = pd.Index(times[offs[members[0]]:offs[members[0]] + c]) if times is not None else pd.RangeIndex(c)
        out.append((index, blk, np.asarray(members)))
    return out, list(uniques)


def process_batch_input(data, pre_period, post_period, series_col: Optional[str] = None,
                        time_col: Optional[str] = None) -> Dict[str, Any]:
    """data: array [N, T, 1 + k]  |  sequence of N DataFrames (column 0 = response)  |  long-format DataFrame with
    `series_col` (and `time_col`).  Returns {'groups': [SeriesGroup...], 'n': N, 'labels': series labels}."""
    if data is None or pre_period is None or post_period is None:
        missing = sorted(k for k, v in {'data': data, 'pre_period': pre_period, 'post_period': post_period}.items()
                         if v is None)
        raise ValueError(f'{", ".join(missing)} input argument{"s" if len(missing) > 1 else ""} cannot be empty')
    labels = None
    raw = []                                                  # (index, block [n, T, 1 + k], positions)
    if isinstance(data, pd.DataFrame):
        if series_col is None:
            raise ValueError('a long-format DataFrame needs `series_col` (use CausalImpact(...) for one series).')
        raw, labels = _from_long(data, series_col, time_col)
    elif isinstance(data, np.ndarray) or hasattr(data, '__array_interface__'):
        arr = np.asarray(data, dtype=np.float32)
        if arr.ndim != 3 or arr.shape[0] < 1:
            raise ValueError('data must be an array of shape [N, T, 1 + k].')
        raw = [(pd.RangeIndex(arr.shape[1]), arr, np.arange(arr.shape[0]))]
    else:
        frames = list(data)
        if not frames or not all(isinstance(f, pd.DataFrame) for f in frames):
            raise ValueError('data must be an array [N, T, 1 + k], a sequence of DataFrames or a long-format DataFrame.')
        # group by index: identity first (frames built from one index object), then equality
        by_id: Dict[int, int] = {}
        reps: List[pd.Index] = []
        members: List[List[int]] = []
        for i, f in enumerate(frames):
            g = by_id.get(id(f.index))
            if g is None:
                for j, rep in enumerate(reps):
                    if len(rep) == len(f.index) and rep.equals(f.index):
                        g = j
                        break
                if g is None:
                    g = len(reps)
                    reps.append(f.index)
                    members.append([])
                by_id[id(f.index)] = g
            members[g].append(i)
        for rep, mem in zip(reps, members):
            raw.append((rep, _to_block([frames[i] for i in mem], mem), np.asarray(mem)))
    groups = [_build_group(block, index, positions, pre_period, post_period) for index, block, positions in raw]
    n = int(sum(len(g.positions) for g in groups))
    return {'groups': groups, 'n': n, 'labels': labels}
