"""Order statistics for the normalisation optimiser (PD:981-1199, PD:1250-1421).

``DeviceOrderStats`` answers rank queries over float32 device volumes with a three-pass
radix select built on ``m3d_select_hist`` (11 + 11 + 10 key bits); the volumes are never
sorted, copied or brought to the host.  The interpolation / median arithmetic on the two
or four selected values follows NumPy exactly (``np.percentile`` ``linear`` method,
``np.median``), which is what the oracle calls.
"""

from __future__ import annotations

import numpy as np

PRED_ALL, PRED_LT, PRED_GT = 0, 1, 2
_SHIFTS = (21, 10, 0)
_WIDTHS = (11, 11, 10)


def _f32_at_or_above(c: float) -> np.float32:
    """smallest float32 >= c   (so that  v < c  <=>  v < result  for float32 v)."""
    f = np.float32(c)
    if float(f) < c:
        f = np.nextafter(f, np.float32(np.inf))
    return f


def _f32_at_or_below(c: float) -> np.float32:
    """largest float32 <= c    (so that  v > c  <=>  v > result  for float32 v)."""
    f = np.float32(c)
    if float(f) > c:
        f = np.nextafter(f, np.float32(-np.inf))
    return f


def _key_to_f32(key: int) -> np.float32:
    u = np.uint32(key)
    u = (u & np.uint32(0x7FFFFFFF)) if (int(u) & 0x80000000) else ~u
    return np.array([u], dtype=np.uint32).view(np.float32)[0]


class DeviceOrderStats:
    """Rank selection over ``v = clip0 ? max(x - sub, 0) : x - sub`` of several volumes."""

    def __init__(self, ctx, reduce=None):
        """``reduce(hist)``: optional in-place sum of the 2048-bin int64 device histogram over the ranks of a
        process group (``dist.all_reduce``).  With it the pooled multiset is the union of every rank's volumes:
        all ranks walk the same digits and end with the same exact order statistics, so the tiles behind a
        pooled median can be sharded over GPUs.  Every rank must issue the same sequence of queries (ranks
        holding no volume pass an empty list)."""
        import torch

        self._ctx = ctx
        self._torch = torch
        self._hist = torch.zeros(2048, dtype=torch.int64, device=ctx.device)
        self._reduce = reduce

    def _pass(self, volumes, sub, clip0, pred, cutoffs, prefix_mask, prefix_value, shift):
        self._hist.zero_()
        for vol, cut in zip(volumes, cutoffs):
            self._ctx.select_hist(
                vol, self._hist, sub=sub, clip0=clip0, pred=pred, cutoff=cut,
                prefix_mask=prefix_mask, prefix_value=prefix_value, shift=shift,
            )
        if self._reduce is not None:
            self._reduce(self._hist)
        return self._hist.cpu().numpy()

    def first_digit_histogram(self, volumes, sub=0.0, clip0=False, pred=PRED_ALL, cutoffs=None):
        """Histogram of the first digit over the whole (predicate-filtered) population: its sum is the population
        size and it is the first level of any ``select`` over the same population."""
        cutoffs = [0.0] * len(volumes) if cutoffs is None else cutoffs
        return self._pass(volumes, sub, clip0, pred, cutoffs, 0, 0, _SHIFTS[0])

    def count(self, volumes, sub=0.0, clip0=False, pred=PRED_ALL, cutoffs=None) -> int:
        return int(self.first_digit_histogram(volumes, sub, clip0, pred, cutoffs).sum())

    def select(self, volumes, ranks, sub=0.0, clip0=False, pred=PRED_ALL, cutoffs=None, first_hist=None):
        """Values at the given 0-based ranks of the pooled, predicate-filtered multiset.  The ranks are walked
        digit by digit TOGETHER: ranks that still share a key prefix (the two middle ranks of a median, the two
        neighbours of a percentile) share that level's histogram pass over the volumes.  ``first_hist``: the
        population's first-digit histogram when the caller already has it (the count behind a median)."""
        cutoffs = [0.0] * len(volumes) if cutoffs is None else cutoffs
        state = {int(r): [0, 0, int(r)] for r in ranks}  # rank -> [prefix_mask, prefix_value, remaining]
        for shift, width in zip(_SHIFTS, _WIDTHS):
            hists = {} if (first_hist is None or shift != _SHIFTS[0]) else {(0, 0): first_hist}
            for st in state.values():
                key = (st[0], st[1])
                if key not in hists:
                    hists[key] = self._pass(volumes, sub, clip0, pred, cutoffs, st[0], st[1], shift)
                h = hists[key]
                # the last digit (10 bits) is histogrammed with the kernel's fixed 11-bit mask: its
                # top bit repeats the previous digit's lowest bit, already pinned by the prefix
                cum = np.cumsum(h)
                b = int(np.searchsorted(cum, st[2], side="right"))
                if b >= h.size:
                    raise ValueError("rank beyond the population")
                st[2] -= int(cum[b - 1]) if b > 0 else 0
                st[1] |= b << shift
                st[0] |= (((1 << width) - 1) << shift) & 0xFFFFFFFF
        return [_key_to_f32(state[int(r)][1]) for r in ranks]

    # ------------------------------------------------------------------ NumPy-exact wrappers
    def percentile(self, volume, q: float, sub=0.0, clip0=False):
        """``np.percentile(v.ravel(), q)`` for a float32 volume, method 'linear'.

        NumPy (2.x) keeps the whole computation in the array's dtype: q/100, the virtual index
        (n-1)*q and the interpolation weight are float32 -- for 4e8-voxel volumes the index is
        therefore quantised, and so it is here.  Returns np.float32 like NumPy."""
        n = int(volume.numel())
        if n == 0:
            return None
        q32 = np.true_divide(q, np.float32(100))
        virtual = np.float32((n - 1) * q32)
        prev = int(np.floor(virtual))
        nxt = prev + 1
        if virtual >= n - 1:
            prev = nxt = n - 1
        if virtual < 0:
            prev = nxt = 0
        a, b = (np.float32(v) for v in self.select([volume], [prev, nxt], sub=sub, clip0=clip0))
        gamma = np.float32(np.float64(virtual) - prev)
        diff = np.float32(b - a)
        if gamma >= np.float32(0.5):
            return np.float32(b - np.float32(diff * np.float32(1 - gamma)))
        return np.float32(a + np.float32(diff * gamma))

    def median(self, volumes, sub=0.0, clip0=False, pred=PRED_ALL, cutoffs=None):
        """``np.median`` of the pooled selected float32 values; None when empty."""
        h0 = self.first_digit_histogram(volumes, sub, clip0, pred, cutoffs)  # one pass: the count AND the first level
        n = int(h0.sum())
        if n == 0:
            return None
        if n % 2 == 1:
            (v,) = self.select(volumes, [n // 2], sub, clip0, pred, cutoffs, first_hist=h0)
            return np.float32(v)
        a, b = self.select(volumes, [n // 2 - 1, n // 2], sub, clip0, pred, cutoffs, first_hist=h0)
        # np.median -> np.mean of the two middle float32 values (float32 add, then / 2)
        return np.float32(np.float32(a) + np.float32(b)) / np.float32(2.0)


def global_normalization_vectors(ctx, bit_volume_lists, low_percentile_cut=10.0,
                                 high_percentile_cut=90.0, reduce=None):
    """PD:1113-1183 -- per bit: bkg = median(pooled pixels < P10_t), nrm = median(pooled
    clip(img - bkg, 0) > P90_t).  ``bit_volume_lists[b]`` = list of float32 device volumes (one
    per sampled tile), already hot-pixel-corrected, z-cropped and low-passed.

    ``reduce`` (see :class:`DeviceOrderStats`): the sampled tiles are sharded over ranks.  The percentile
    cut-offs are per tile (local); the two medians are over the pool of all ranks' selected pixels, found by
    the same radix select with the digit histograms summed across ranks -- exact, so the vectors equal the
    single-process ones bit for bit.  A rank without tiles passes empty lists and still takes part."""
    local = DeviceOrderStats(ctx)
    pooled = local if reduce is None else DeviceOrderStats(ctx, reduce=reduce)
    n_bits = len(bit_volume_lists)
    nrm = np.ones(n_bits, dtype=np.float32)
    bkg = np.zeros(n_bits, dtype=np.float32)
    for b, vols in enumerate(bit_volume_lists):
        vols = [v for v in vols if v.numel() > 0]
        if not vols and reduce is None:
            continue
        cuts = [float(local.percentile(v, low_percentile_cut)) for v in vols]  # float32 cut-offs, like NumPy
        m = pooled.median(vols, pred=PRED_LT, cutoffs=cuts)
        bkg[b] = 0 if m is None else m
        sub = float(bkg[b])
        cuts = [float(local.percentile(v, high_percentile_cut, sub=sub, clip0=True)) for v in vols]
        m = pooled.median(vols, sub=sub, clip0=True, pred=PRED_GT, cutoffs=cuts)
        nrm[b] = 1 if m is None else m
    return nrm, bkg


def iterative_normalization_vectors(df, n_bits: int):
    """PD:1263-1368 vectorised: per-bit medians of on-bit / off-bit feature means over
    non-blank transcripts, cast float32, rounded to 0.1, NaN -> 1/0, 0 -> 1.

    Returns ``(normalization, background)`` or ``None`` when the reference keeps the
    previous vectors (no non-blank transcripts or no bit columns).
    """
    import pandas as pd  # noqa: F401  (df is a DataFrame)

    keep = ~df["gene_id"].astype("string").str.lower().str.startswith("blank", na=False)
    d = df[keep]
    bit_cols = [c for c in d.columns if c.startswith("bit") and c.endswith("_mean_intensity")]
    if d.empty or not bit_cols:
        return None
    # the reference sorts the column names (PD:1338-1343) before the median
    bit_cols = sorted(bit_cols)
    vals = d[bit_cols].to_numpy(dtype=np.float64)
    col_bit = np.array([int(c[3:5]) for c in bit_cols])
    on = d[[f"on_bit_{k}" for k in range(1, 5)]].to_numpy(dtype=np.int64)
    is_on = (on[:, :, None] == col_bit[None, None, :]).any(axis=1)
    # per-bit medians of the selected entries (pandas ``median(skipna=True)`` of the reference's sparse
    # frames, PD:1345-1356): NaN = no entry
    finite = ~np.isnan(vals)
    med_on = np.full(len(bit_cols), np.nan)
    med_off = np.full(len(bit_cols), np.nan)
    for j in range(len(bit_cols)):
        sel_on = vals[is_on[:, j] & finite[:, j], j]
        sel_off = vals[~is_on[:, j] & finite[:, j], j]
        if sel_on.size:
            med_on[j] = np.median(sel_on)
        if sel_off.size:
            med_off[j] = np.median(sel_off)
    nv = np.round(med_on.astype(np.float32), 1)
    bv = np.round(med_off.astype(np.float32), 1)
    nv = np.nan_to_num(nv, 1.0)
    nv = np.where(nv == 0.0, 1.0, nv)
    bv = np.nan_to_num(bv, 0.0)
    return nv.astype(np.float32), bv.astype(np.float32)


def _numpy_hist_backend(data, hist_row, prefix_mask, prefix_value, shift):
    """Host restatement of ``select_hist_kernel`` (csrc/api.cu) for tensors that live on the CPU: the digit
    histogram of the order-preserving uint32 keys.  Used by the CPU tests of the multi-rank host logic."""
    u = data.detach().cpu().numpy().astype(np.float32).view(np.uint32)
    key = np.where(u & np.uint32(0x80000000), ~u, u | np.uint32(0x80000000)).astype(np.uint32)
    sel = (key & np.uint32(prefix_mask)) == np.uint32(prefix_value)
    digits = (key[sel] >> np.uint32(shift)) & np.uint32(2047)
    h = np.bincount(digits.astype(np.int64), minlength=2048).astype(np.int64)
    import torch

    hist_row += torch.from_numpy(h)


def pooled_medians(queries, hist_fn, new_hist, reduce=None, hist_batch_fn=None):
    """``np.median`` of each query's values pooled over all ranks, without gathering them.

    ``queries``: list of 1-D float32 tensors (this rank's part of every multiset; may be empty).
    ``hist_fn(data, hist_row, prefix_mask, prefix_value, shift)`` adds the 2048-bin digit histogram of ``data``'s
    keys into ``hist_row`` (``DecodeContext.select_hist`` on the device); ``new_hist(rows)`` allocates a zeroed
    ``(rows, 2048)`` int64 tensor; ``reduce(hist)`` sums it in place over the process group (``all_reduce``).
    ``hist_batch_fn(rows, hist, shift)`` (optional; ``DecodeContext.select_hist_batch``) fills all rows of one round with
    ONE launch -- ``rows`` = list of ``(data, prefix_mask, prefix_value)`` -- instead of one ``hist_fn`` call per row.

    Three rounds, each ONE collective and ONE device-to-host copy for all queries together: the first digit's
    histograms (which also give the pooled counts), then the second and third digit of the one or two middle
    ranks of every query.  Exact: the result is the float64 mean of the two middle float32 values, like
    ``np.median`` of the float64 column the reference builds (PD:1345-1356).  Returns a list of float (NaN where
    the pooled multiset is empty) -- identical on every rank."""
    nq = len(queries)
    out = [float("nan")] * nq
    if nq == 0:
        return out

    def run(rows):
        """rows: list of (query index, prefix_mask, prefix_value, shift) -> (len(rows), 2048) host histograms"""
        hist = new_hist(len(rows))
        if hist_batch_fn is not None and rows:
            hist_batch_fn([(queries[q], pm, pv) for q, pm, pv, _sh in rows], hist, rows[0][3])  # one shift per round
        else:
            for r, (q, pm, pv, sh) in enumerate(rows):
                if queries[q].numel():
                    hist_fn(queries[q], hist[r], pm, pv, sh)
        if reduce is not None:
            reduce(hist)
        return hist.cpu().numpy()

    first = run([(q, 0, 0, _SHIFTS[0]) for q in range(nq)])
    counts = first.sum(axis=1)
    # targets: (query, rank) with per-target walking state [prefix_mask, prefix_value, remaining]
    targets = {}
    for q in range(nq):
        n = int(counts[q])
        if n == 0:
            continue
        for r in sorted({(n - 1) // 2, n // 2}):
            targets[(q, r)] = [0, 0, r]
    for level, (shift, width) in enumerate(zip(_SHIFTS, _WIDTHS)):
        if level == 0:
            hist_of = {(q, 0, 0): first[q] for q in range(nq)}
        else:
            keys = sorted({(q, st[0], st[1]) for (q, _r), st in targets.items()})
            hs = run([(q, pm, pv, shift) for q, pm, pv in keys]) if keys else np.zeros((0, 2048), dtype=np.int64)
            hist_of = {k: hs[i] for i, k in enumerate(keys)}
        for (q, _r), st in targets.items():
            h = hist_of[(q, st[0], st[1])]
            cum = np.cumsum(h)
            b = int(np.searchsorted(cum, st[2], side="right"))
            if b >= h.size:
                raise ValueError("rank beyond the population")
            st[2] -= int(cum[b - 1]) if b > 0 else 0
            st[1] |= b << shift
            st[0] |= (((1 << width) - 1) << shift) & 0xFFFFFFFF
    for q in range(nq):
        n = int(counts[q])
        if n == 0:
            continue
        a = np.float64(_key_to_f32(targets[(q, (n - 1) // 2)][1]))
        b = np.float64(_key_to_f32(targets[(q, n // 2)][1]))
        out[q] = float((a + b) / 2.0) if (n % 2 == 0) else float(a)
    return out


def finish_iterative_vectors(med_on, med_off):
    """PD:1358-1368: float32, round to 0.1, NaN -> 1 / 0, normalisation 0 -> 1."""
    nv = np.round(np.asarray(med_on, dtype=np.float64).astype(np.float32), 1)
    bv = np.round(np.asarray(med_off, dtype=np.float64).astype(np.float32), 1)
    nv = np.nan_to_num(nv, 1.0)
    nv = np.where(nv == 0.0, 1.0, nv)
    bv = np.nan_to_num(bv, 0.0)
    return nv.astype(np.float32), bv.astype(np.float32)


def non_blank_rows(df) -> np.ndarray:
    """Boolean mask of the rows whose gene id does not start with 'blank' (case-insensitive; missing ids are kept,
    PD:1263-1270) -- evaluated once per distinct id instead of once per row."""
    import pandas as pd

    codes, uniques = pd.factorize(df["gene_id"], use_na_sentinel=True)
    blank_u = np.array([str(u).lower().startswith("blank") for u in uniques], dtype=bool)
    if blank_u.size == 0:
        return np.ones(len(df), dtype=bool)
    return np.where(codes >= 0, ~blank_u[np.maximum(codes, 0)], True)


def iterative_vector_queries(df, n_bits: int, device):
    """The 2 x bits multisets behind the iterative vectors (PD:1290-1356) as float32 tensors on ``device``:
    for every bit column (sorted by name like the reference) the mean intensities of the non-blank rows that have
    the bit ON, then those that have it OFF.  Returns ``(queries, n_rows_kept)`` or ``(None, 0)`` when the table has
    no bit columns; values that are not exactly float32-representable or not finite make it return ``None`` too
    (foreign tables: the caller falls back to the host median)."""
    import torch

    bit_cols = sorted(c for c in df.columns if c.startswith("bit") and c.endswith("_mean_intensity"))
    if not bit_cols:
        return None, 0
    keep = non_blank_rows(df) if len(df) else np.zeros(0, dtype=bool)
    vals = df[bit_cols].to_numpy(dtype=np.float64)[keep]
    on = df[[f"on_bit_{k}" for k in range(1, 5)]].to_numpy(dtype=np.int64)[keep]
    v64 = torch.from_numpy(np.ascontiguousarray(vals)).to(device)
    v32 = v64.to(torch.float32)
    if v64.numel() and not bool(((v32.to(torch.float64) == v64) | torch.isnan(v64)).all()):
        return None, int(keep.sum())
    col_bit = [int(c[3:5]) for c in bit_cols]
    col_of_bit = torch.full((max(col_bit + [int(on.max()) if on.size else 0]) + 2,), len(bit_cols), dtype=torch.int64,
                            device=device)
    col_of_bit[torch.tensor(col_bit, dtype=torch.int64, device=device)] = torch.arange(len(bit_cols), device=device)
    is_on = torch.zeros((v32.shape[0], len(bit_cols) + 1), dtype=torch.bool, device=device)
    if on.size:
        on_d = torch.from_numpy(np.ascontiguousarray(on)).to(device).clamp_(min=0)
        is_on.scatter_(1, col_of_bit[on_d], True)
    is_on = is_on[:, : len(bit_cols)]
    finite = ~torch.isnan(v32)
    queries = []
    for j in range(len(bit_cols)):
        col = v32[:, j]
        queries.append(col[is_on[:, j] & finite[:, j]].contiguous())
    for j in range(len(bit_cols)):
        col = v32[:, j]
        queries.append(col[~is_on[:, j] & finite[:, j]].contiguous())
    return queries, int(keep.sum())
