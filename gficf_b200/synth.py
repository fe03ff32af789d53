"""Synthetic kNN index matrices of the shapes BASELINE.json names (there is no dataset here).

Counter-based (splitmix64 of (seed, row, slot)) integer arithmetic in torch, so the same call
gives bit-identical matrices on CPU and on a GPU.  Rows hold k DISTINCT ids, none equal to the
row itself -- what ``neigh[,-1]`` of an Annoy kNN result looks like (R/clustCells.R:60-63).

Families:
  planted   clusters of `cluster` consecutive cells; each neighbour from the own cluster with
            probability p_in, else uniform over all cells
  uniform   k ids uniform without replacement (worst locality; u is almost always 0)
`scramble=True` relabels cells with the bijection x -> (a*x + b) mod n (applied to row positions
and to values) so that neighbouring rows are no longer neighbouring in memory.
"""
from __future__ import annotations

import math

import torch

_M64 = (1 << 64) - 1


def _i64(x: int) -> int:
    x &= _M64
    return x - (1 << 64) if x >= (1 << 63) else x


_C1, _C2, _C3 = _i64(0x9E3779B97F4A7C15), _i64(0xBF58476D1CE4E5B9), _i64(0x94D049BB133111EB)


def _srl(x: torch.Tensor, s: int) -> torch.Tensor:
    """Logical right shift of int64 (torch's >> is arithmetic)."""
    return (x >> s) & ((1 << (64 - s)) - 1)


def splitmix64(x: torch.Tensor) -> torch.Tensor:
    z = x + _C1
    z = (z ^ _srl(z, 30)) * _C2
    z = (z ^ _srl(z, 27)) * _C3
    return z ^ _srl(z, 31)


def _uniform_bits(seed: int, rows: torch.Tensor, slots: torch.Tensor, stream: int) -> torch.Tensor:
    """63 random bits per (row, slot)."""
    key = splitmix64(rows * _i64(0xD1342543DE82EF95) + _i64(seed * 0x2545F4914F6CDD1D + stream))
    return _srl(splitmix64(key[:, None] ^ (slots[None, :] * _i64(0xA0761D6478BD642F))), 1)


def _first_k_distinct(cand: torch.Tensor, self_id: torch.Tensor, k: int):
    """Per row: the first k candidates that are distinct and != self, in candidate order.
    Returns (idx [n,k], ok [n])."""
    n, m = cand.shape
    srt, order = torch.sort(cand, dim=1, stable=True)
    dup_sorted = torch.zeros_like(cand, dtype=torch.bool)
    dup_sorted[:, 1:] = srt[:, 1:] == srt[:, :-1]
    bad = torch.zeros_like(dup_sorted)
    bad.scatter_(1, order, dup_sorted)
    bad |= cand == self_id[:, None]
    good = ~bad
    rank = torch.cumsum(good, dim=1)
    ok = rank[:, -1] >= k
    take = good & (rank <= k)
    # positions of the taken candidates, in order: stable sort of ~take puts them first
    pos = torch.sort((~take).to(torch.int8), dim=1, stable=True).indices[:, :k]
    return torch.gather(cand, 1, pos), ok


def knn_index(n: int, k: int, family: str = "planted", seed: int = 180582, scramble: bool = False,
              cluster: int | None = None, p_in: float = 0.95, device="cpu", chunk: int = 1 << 20) -> torch.Tensor:
    """int32 [n, k], 0-based neighbour ids."""
    if n - 1 < k:
        raise ValueError("need n > k")
    if cluster is None:
        cluster = 256 if k <= 30 else 1024
    cluster = min(cluster, n)
    m = 2 * k + 8
    out = torch.empty((n, k), dtype=torch.int32, device=device)
    slots = torch.arange(m, dtype=torch.int64, device=device)
    thresh = int(p_in * (1 << 62))
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        rows = torch.arange(lo, hi, dtype=torch.int64, device=device)
        r1 = _uniform_bits(seed, rows, slots, 1)
        if family == "planted":
            r2 = _uniform_bits(seed, rows, slots, 2)
            base = (rows // cluster) * cluster
            size = torch.clamp(torch.full_like(base, cluster), max=n) 
            size = torch.minimum(size, n - base)
            local = base[:, None] + r1 % size[:, None]
            cand = torch.where(_srl(r2, 1) < thresh, local, r1 % n)
        elif family == "uniform":
            cand = r1 % n
        else:
            raise ValueError("family must be 'planted' or 'uniform'")
        idx, ok = _first_k_distinct(cand, rows, k)
        if not bool(ok.all()):
            # rare: top up deficient rows deterministically with the next free ids
            for r in torch.nonzero(~ok).flatten().tolist():
                i = lo + r
                have = []
                for v in cand[r].tolist():
                    if v != i and v not in have:
                        have.append(v)
                v = (i + 1) % n
                while len(have) < k:
                    if v != i and v not in have:
                        have.append(v)
                    v = (v + 1) % n
                idx[r] = torch.tensor(have[:k], dtype=torch.int64, device=device)
        out[lo:hi] = idx.to(torch.int32)
    if scramble:
        out = scramble_ids(out, seed)
    return out


def scramble_ids(idx: torch.Tensor, seed: int = 180582) -> torch.Tensor:
    """Relabel cells with x -> (a*x + b) mod n: row pi(i) of the result is pi(row i of idx)."""
    n = idx.shape[0]
    a, b = _scramble_coefficients(n, seed)
    rows = torch.arange(n, dtype=torch.int64, device=idx.device)
    new_pos = (rows * a + b) % n
    vals = ((idx.to(torch.int64) * a + b) % n).to(torch.int32)
    out = torch.empty_like(idx)
    out[new_pos] = vals
    return out


def _scramble_coefficients(n: int, seed: int):
    a = (0x9E3779B1 * (2 * (seed % 1000) + 1)) % n
    if a < 2:
        a = 1 if n <= 2 else 2
    while math.gcd(a, n) != 1:
        a += 1
    return a, (seed * 7919) % n


def planted_community(n: int, k: int, seed: int = 180582, scramble: bool = False, cluster: int | None = None,
                      device="cpu") -> torch.Tensor:
    """int32 [n]: the planted community of every cell of knn_index(n, k, family="planted", ...), in the ids
    that call returns (the block of `cluster` consecutive ORIGINAL ids the cell belongs to)."""
    if cluster is None:
        cluster = 256 if k <= 30 else 1024
    cluster = min(cluster, n)
    ids = torch.arange(n, dtype=torch.int64, device=device)
    if scramble:
        a, b = _scramble_coefficients(n, seed)
        ids = ((ids - b) % n) * pow(a, -1, n) % n  # inverse of x -> (a*x + b) mod n
    return (ids // cluster).to(torch.int32)


def to_r_matrix(idx0: torch.Tensor):
    """int32 0-based [n,k] -> what R hands to the native code: float64, column-major, 1-based
    (numpy, Fortran order)."""
    import numpy as np

    a = idx0.cpu().numpy()
    return np.asfortranarray(a.astype(np.float64) + 1.0)
