"""Device-side synthetic paintings for the benchmark configurations: the same
generative recipe as :mod:`synth` (2-source ancestry tracts, exponential donor-haplotype runs),
vectorised with torch so that 2,000 paintings x 100k SNPs are produced in HBM in well under a
second.  torch is plumbing here (device memory + RNG); nothing on the measured path uses it."""
from __future__ import annotations

import numpy as np
import torch

from .synth import SynthSpec, recomb_map, cm_positions


@torch.no_grad()
def make_chromosome_device(spec: SynthSpec, chrom: int, device, ind_offset: int = 0, rows_per_batch: int = 400):
    """Returns (pos, rate) numpy float64 and labels: torch.int16 [n_inds*ploidy*nsamples, n_snps] on
    ``device`` (values < 32768, bit-identical to uint16).  ``ind_offset`` decorrelates shards."""
    assert spec.n_donor_haps < 32768
    pos, rate = recomb_map(spec, chrom)
    cm = torch.from_numpy(cm_positions(pos, rate)).to(device)
    total = float(cm[-1])
    g = torch.Generator(device=device)
    g.manual_seed(spec.seed + 104729 * (chrom + 1) + 7919 * ind_offset)
    prng = np.random.default_rng(spec.seed + 1)
    srcvec = torch.from_numpy(prng.dirichlet(np.full(spec.K, 0.6), size=2)).to(device)
    cdf = torch.cumsum(srcvec, dim=1).contiguous()
    rows = spec.n_inds * spec.ploidy * spec.nsamples
    L = spec.n_snps
    labels = torch.empty((rows, L), dtype=torch.int16, device=device)
    nc = int(total / spec.mean_chunk_cm * 1.35) + 64
    nt = int(total * spec.gens / 100.0 * 1.6) + 16
    for r0 in range(0, rows, rows_per_batch):
        r1 = min(rows, r0 + rows_per_batch)
        R = r1 - r0
        lens = -torch.log1p(-torch.rand((R, nc), generator=g, device=device, dtype=torch.float64)) * spec.mean_chunk_cm
        cb = torch.cumsum(lens, dim=1)
        cb = torch.cat([torch.zeros((R, 1), dtype=torch.float64, device=device), cb[:, :-1]], dim=1)   # chunk starts
        tl = -torch.log1p(-torch.rand((R, nt), generator=g, device=device, dtype=torch.float64)) * (100.0 / spec.gens)
        tb = torch.cumsum(tl, dim=1)
        src0 = torch.randint(0, 2, (R, 1), generator=g, device=device)
        src = (src0 + torch.searchsorted(tb, cb, right=True)) & 1
        u = torch.rand((R, nc), generator=g, device=device, dtype=torch.float64)
        pop0 = torch.searchsorted(cdf[0].expand(R, -1).contiguous(), u)
        pop1 = torch.searchsorted(cdf[1].expand(R, -1).contiguous(), u)
        pop = torch.where(src == 0, pop0, pop1).clamp_(0, spec.K - 1)
        hap = pop * spec.haps_per_pop + torch.randint(0, spec.haps_per_pop, (R, nc), generator=g, device=device) + 1
        # chunk index of every SNP: number of chunk starts <= cM position, minus one
        idx = torch.searchsorted(cb, cm.expand(R, -1).contiguous(), right=True) - 1
        idx.clamp_(0, nc - 1)
        labels[r0:r1] = torch.gather(hap, 1, idx).to(torch.int16)
        del lens, cb, tl, tb, src, u, pop0, pop1, pop, hap, idx
    return pos, rate, labels


@torch.no_grad()
def make_chromosome_runs_device(spec: SynthSpec, chrom: int, device, ind_offset: int = 0, rows_per_batch: int = 400):
    """The same paintings as :func:`make_chromosome_device` (same random stream), produced directly in the library's
    run-length-encoded form (``fgt_run_t``: first SNP, donor haplotype) without ever materialising the dense label matrix --
    which is what makes the biobank-scale configurations (1M SNPs per chromosome) fit.
    Returns (pos, rate) numpy float64, run_off numpy int64 [rows+1] and runs: torch.int32 [n_runs, 2] on ``device``."""
    pos, rate = recomb_map(spec, chrom)
    cm = torch.from_numpy(cm_positions(pos, rate)).to(device)
    total = float(cm[-1])
    g = torch.Generator(device=device)
    g.manual_seed(spec.seed + 104729 * (chrom + 1) + 7919 * ind_offset)
    prng = np.random.default_rng(spec.seed + 1)
    srcvec = torch.from_numpy(prng.dirichlet(np.full(spec.K, 0.6), size=2)).to(device)
    cdf = torch.cumsum(srcvec, dim=1).contiguous()
    rows = spec.n_inds * spec.ploidy * spec.nsamples
    L = spec.n_snps
    nc = int(total / spec.mean_chunk_cm * 1.35) + 64
    nt = int(total * spec.gens / 100.0 * 1.6) + 16
    pieces, counts = [], []
    for r0 in range(0, rows, rows_per_batch):
        r1 = min(rows, r0 + rows_per_batch)
        R = r1 - r0
        lens = -torch.log1p(-torch.rand((R, nc), generator=g, device=device, dtype=torch.float64)) * spec.mean_chunk_cm
        cb = torch.cumsum(lens, dim=1)
        cb = torch.cat([torch.zeros((R, 1), dtype=torch.float64, device=device), cb[:, :-1]], dim=1)   # chunk starts
        tl = -torch.log1p(-torch.rand((R, nt), generator=g, device=device, dtype=torch.float64)) * (100.0 / spec.gens)
        tb = torch.cumsum(tl, dim=1)
        src0 = torch.randint(0, 2, (R, 1), generator=g, device=device)
        src = (src0 + torch.searchsorted(tb, cb, right=True)) & 1
        u = torch.rand((R, nc), generator=g, device=device, dtype=torch.float64)
        pop0 = torch.searchsorted(cdf[0].expand(R, -1).contiguous(), u)
        pop1 = torch.searchsorted(cdf[1].expand(R, -1).contiguous(), u)
        pop = torch.where(src == 0, pop0, pop1).clamp_(0, spec.K - 1)
        hap = pop * spec.haps_per_pop + torch.randint(0, spec.haps_per_pop, (R, nc), generator=g, device=device) + 1
        # SNP i copies chunk (number of chunk starts <= cm[i]) - 1, so chunk c covers the SNPs from the first one at or
        # behind its start up to the next chunk's first SNP; the last chunk runs to the end of the chromosome
        lo = torch.searchsorted(cm, cb)                                        # first SNP with cm >= start
        lo_next = torch.cat([lo[:, 1:], torch.full((R, 1), L, dtype=lo.dtype, device=device)], dim=1)
        ri, ci = (lo < lo_next).nonzero(as_tuple=True)                         # chunks that hold at least one SNP, row-major
        start, h = lo[ri, ci], hap[ri, ci]
        new = torch.ones_like(ri, dtype=torch.bool)
        new[1:] = (ri[1:] != ri[:-1]) | (h[1:] != h[:-1])                      # maximal runs: merge equal neighbours
        sel = new.nonzero(as_tuple=True)[0]
        pieces.append(torch.stack([start[sel].to(torch.int32), h[sel].to(torch.int32)], dim=1))
        counts.append(torch.bincount(ri[sel], minlength=R).cpu())
        del lens, cb, tl, tb, src, u, pop0, pop1, pop, hap, lo, lo_next, ri, ci, start, h, new, sel
    runs = torch.cat(pieces, dim=0).contiguous()
    cnt = torch.cat(counts).numpy().astype(np.int64)
    run_off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    return pos, rate, run_off, runs


@torch.no_grad()
def rle_device(labels: torch.Tensor):
    """dense [rows, L] label matrix on the device -> (run_off numpy int64 [rows+1], runs torch.int32 [n_runs, 2])"""
    rows, L = labels.shape
    pieces, counts = [], []
    for r0 in range(0, rows, 256):
        lab = labels[r0:r0 + 256]
        change = torch.ones(lab.shape, dtype=torch.bool, device=labels.device)
        change[:, 1:] = lab[:, 1:] != lab[:, :-1]
        ri, ci = change.nonzero(as_tuple=True)
        pieces.append(torch.stack([ci.to(torch.int32), lab[ri, ci].to(torch.int32) & 0xFFFF if lab.dtype == torch.int16 else lab[ri, ci].to(torch.int32)], dim=1))
        counts.append(torch.bincount(ri, minlength=lab.shape[0]).cpu())
    runs = torch.cat(pieces, dim=0).contiguous()
    cnt = torch.cat(counts).numpy().astype(np.int64)
    return np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64), runs
