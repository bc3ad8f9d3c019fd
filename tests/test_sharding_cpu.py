"""world_size-2 gloo test (CPU) of the multi-rank protocol: ranks exchange per-(chromosome, individual)
pair counts with one all-gather, derive their rand() stream offsets, position a glibc-compatible
stream with the library's host jump-ahead, and together reproduce the unsharded raw tensors
bit-for-bit; the all-reduced curves agree to 1e-12."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import ROOT

K, S, G, BW, GS, SEED = 5, 2, 70, 0.1, 10, 321


def _data():
    from fastglobetrotter_b200 import synth
    spec = synth.SynthSpec(n_chrom=1, n_inds=4, n_snps=2000, K=K, nsamples=3, rate=6e-7, seed=17)
    pos, rate, labels = synth.make_chromosome(spec, 0)
    return spec, [{"pos": pos, "rate": rate, "labels": labels}]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fastglobetrotter_b200 import shard
    from fastglobetrotter_b200.companion import PackedCompanion
    from oracle.oracle import Oracle
    lib, orc = PackedCompanion(), Oracle()
    spec, chrom = _data()
    dl = spec.donor_label_vec()
    tw = np.random.default_rng(0).random(K * K * S * S)
    lo, hi = shard.block_range(spec.n_inds, rank, world)
    # local pair counts (one oracle pass per individual; the count does not depend on the draws)
    counts = np.zeros((1, hi - lo), dtype=np.int64)
    for i, n in enumerate(range(lo, hi)):
        orc.srand(1)
        counts[0, i] = orc.data_read_packed(1, spec.ploidy, spec.nsamples, chrom, [n], BW, G, GS, K, dl, 0.0, 1.0, S, tw)[2]
    gathered = [None] * world
    dist.all_gather_object(gathered, counts)
    offs, total = shard.stream_offsets(gathered)
    hist = np.zeros(31, dtype=np.uint32)
    raws, means = [], np.zeros((S * S, G))
    for i, n in enumerate(range(lo, hi)):
        assert lib.lib.fgt_rand_host_jump(C.c_uint(SEED), C.c_uint64(2 * int(offs[rank][0, i])), hist.ctypes.data_as(C.c_void_p)) == 0
        orc.set_history(hist)
        m, raw, _ = orc.data_read_packed(1, spec.ploidy, spec.nsamples, chrom, [n], BW, G, GS, K, dl, 0.0, 1.0, S, tw, want_raw=True)
        raws.append(raw[0])
        means += m / spec.n_inds          # a single-individual call divides by 1; the global mean divides by N
    t = torch.from_numpy(means)
    dist.all_reduce(t)
    allraw = [None] * world
    dist.all_gather_object(allraw, np.stack(raws))
    if rank == 0:
        q.put((np.concatenate(allraw), t.numpy(), total))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_offsets_reproduce_unsharded_run(oracle):
    spec, chrom = _data()
    dl = spec.donor_label_vec()
    tw = np.random.default_rng(0).random(K * K * S * S)
    ids = np.arange(spec.n_inds, dtype=np.int32)
    oracle.srand(SEED)
    m_all, raw_all, pairs_all = oracle.data_read_packed(1, spec.ploidy, spec.nsamples, chrom, ids, BW, G, GS, K, dl, 0.0, 1.0, S,
                                                        tw, want_raw=True)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    raw2, m2, total = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert total == pairs_all
    assert np.array_equal(raw2, raw_all)
    assert np.allclose(m2, m_all, rtol=1e-12, atol=0)


def test_block_ranges_cover_everything():
    from fastglobetrotter_b200 import shard
    for n in (1, 7, 100, 1001):
        for w in (1, 2, 3, 8):
            cover = []
            for r in range(w):
                lo, hi = shard.block_range(n, r, w)
                cover += list(range(lo, hi))
            assert cover == list(range(n))
