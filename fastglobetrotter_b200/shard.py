"""Sharding of target individuals over ranks.

Every rank owns a contiguous block of pseudo-individuals and all their chromosomes.  The only
cross-rank quantity the sampled-pair path needs is where each (chromosome, individual) starts in
the shared rand() stream: the reference draws in chromosome-major, individual-minor order
(reference C:338, C:419), so the start is the exclusive prefix sum of the per-(h, n) pair counts in
that order.  Counts are exchanged with ONE all-gather; the S^2 x G partial curves are summed with
ONE all-reduce.
"""
from __future__ import annotations

import numpy as np


def block_range(n_total: int, rank: int, world: int):
    """contiguous block [lo, hi) of individuals owned by `rank`"""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def stream_offsets(counts_by_rank):
    """counts_by_rank: list over ranks of int64 arrays [n_chrom, n_local_r] (pairs each (h, n) samples).
    Returns (offsets_by_rank: list of [n_chrom, n_local_r] absolute pair offsets, total_pairs)."""
    flat = np.concatenate([np.asarray(c, dtype=np.int64) for c in counts_by_rank], axis=1)     # [n_chrom, N_total]
    offs = np.concatenate([[0], np.cumsum(flat.ravel())])[:-1].reshape(flat.shape)
    out, lo = [], 0
    for c in counts_by_rank:
        n = np.asarray(c).shape[1]
        out.append(np.ascontiguousarray(offs[:, lo:lo + n]))
        lo += n
    return out, int(flat.sum())
