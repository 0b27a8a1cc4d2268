"""A CPU model of the GPU path's EXACT QUANTILE WITHOUT SORTING (staticmapping_b200/csrc/icp_dev.cuh dist_bin /
sub_bin / select_bin, icp.cu phase B, icp_finish.cu phase C) — TEST INFRASTRUCTURE ONLY.

The reference trims at `limit = nth_element(values, int(size * (double)0.7f))` and keeps `d2 <= limit`
(registrators/icp_fast.cc:65-90, :497-498).  The kernels never sort: phase A histograms d2 by (exponent, 5 mantissa
bits) into 2048 bins, phase B locates the bin that holds the rank, takes everything in lower bins and sub-histograms
the bin's members by the next 11 bits, phase C takes the lower sub-bins, ranks the few keys of one sub-bin by
counting — and falls back to an 8-pass MSB radix select over the bin's members when the bin is a clamp bin (all
d2 == 0, or beyond 2^-40..2^24) or a sub-bin holds more than 1024 keys.  DESIGN.md claims the limit is the value
std::nth_element returns and the kept set is the reference's; this model restates the selection so that claim is
checked on a CPU against numpy's partition, including the fallback paths."""
from __future__ import annotations

import numpy as np

K_HIST_BINS = 2048
K_MAX_EXACT_KEYS = 1024


def _bits(d2):
    return np.asarray(d2, dtype=np.float64).view(np.int64)


def dist_bin(d2):
    key = (_bits(d2) >> 47) - ((1023 - 40) << 5)
    return np.clip(key, 0, K_HIST_BINS - 1)


def sub_bin(d2):
    return (_bits(d2) >> 36) & 2047


def select(d2_all, ratio_f32=0.7):
    """-> dict(limit, kept mask over d2_all, path) following phases A-C; +inf entries are 'no match'."""
    d2_all = np.asarray(d2_all, dtype=np.float64)
    finite = d2_all < np.inf                                   # finite_d2 (NaN is not < inf either)
    vals = d2_all[finite]
    total = vals.size
    if total == 0:
        return {"limit": None, "kept": np.zeros(d2_all.size, bool), "path": "empty"}
    # phase A: global histogram; select_bin: quantile index and the bin that holds it
    hist = np.bincount(dist_bin(vals), minlength=K_HIST_BINS)
    q = float(np.float32(ratio_f32))
    qi = total - 1 if q == 1.0 else int(total * q)
    qi = min(qi, total - 1)
    excl = np.concatenate([[0], np.cumsum(hist)[:-1]])
    b = int(np.flatnonzero((excl <= qi) & (qi < excl + hist))[0])
    below = int(excl[b])
    r = qi - below                                             # rank inside the quantile bin
    bins_all = np.where(finite, dist_bin(np.where(finite, d2_all, 0.0)), -1)
    kept = finite & (bins_all < b)                             # phase B: strictly lower bins are members outright
    cand = np.flatnonzero(finite & (bins_all == b))            # the bin's members, ascending point index
    keys = _bits(d2_all[cand]).astype(np.uint64)
    fallback = b <= 0 or b >= K_HIST_BINS - 1                  # clamp_bin
    path = "two-level"
    if not fallback:
        h2 = np.bincount(sub_bin(d2_all[cand]), minlength=2048)
        e2 = np.concatenate([[0], np.cumsum(h2)[:-1]])
        sb = int(np.flatnonzero((e2 <= r) & (r < e2 + h2))[0])
        if h2[sb] > K_MAX_EXACT_KEYS:
            fallback = True
    if not fallback:
        subs = sub_bin(d2_all[cand])
        kept[cand[subs < sb]] = True                           # lower sub-bins
        sel = np.flatnonzero(subs == sb)
        r2 = r - int(e2[sb])
        # rank by counting, ties broken by slot (= position in the candidate list)
        k = keys[sel]
        pos = np.array([int(np.sum((k < k[i]) | ((k == k[i]) & (np.arange(k.size) < i)))) for i in range(k.size)])
        skeys = np.empty_like(k); skeys[pos] = k
        limit_bits = skeys[r2]
        kept[cand[sel[k <= limit_bits]]] = True
    else:
        path = "radix-select"
        rank, prefix = r, np.uint64(0)
        for p in range(7, -1, -1):
            shift = np.uint64(p * 8)
            mask = np.uint64(0) if p == 7 else (~np.uint64(0)) << np.uint64(p * 8 + 8)
            m = (keys & mask) == prefix
            digits = ((keys[m] >> shift) & np.uint64(255)).astype(np.int64)
            h = np.bincount(digits, minlength=256)
            e = np.concatenate([[0], np.cumsum(h)[:-1]])
            d = int(np.flatnonzero((e <= rank) & (rank < e + h))[0])
            rank -= int(e[d])
            prefix = prefix | (np.uint64(d) << shift)
        limit_bits = prefix
        limit = np.array([limit_bits], dtype=np.uint64).view(np.float64)[0]
        kept[cand[d2_all[cand] <= limit]] = True
    limit = np.array([limit_bits], dtype=np.uint64).view(np.float64)[0]
    return {"limit": float(limit), "kept": kept, "path": path}
