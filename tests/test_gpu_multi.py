"""Multi-GPU inside the library, on real devices: needs >= 2 GPUs (gpurun --gpus 2).

* single process, "devices" = 2: the R-facing call spreads over two GPUs (one engine + host thread per device, NCCL
  all-gather of pair counts, device-side stream offsets, NCCL all-reduce of the curves): raw tensors bit-exact, curves
  <= 1e-12 relative, generator state advanced like the reference's.
* two processes, one GPU each (the torchrun topology of bench.py): fgt_dist_unique_id / fgt_dist_init.
"""
import ctypes as C
import multiprocessing as mp
import os
import sys

import numpy as np
import pytest

from helpers import bootstrap_ids, ind_id_matrix
from fastglobetrotter_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


needs2 = pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")


def _problem(n_inds=7, n_chrom=3):
    spec = synth.SynthSpec(n_chrom=n_chrom, n_inds=n_inds, n_snps=3500, K=6, nsamples=4, rate=4e-7, self_copy_frac=0.03, seed=99, name="multi")
    chrom = []
    for h in range(spec.n_chrom):
        pos, rate, labels = synth.make_chromosome(spec, h)
        chrom.append({"pos": pos, "rate": rate, "labels": labels})
    return spec, chrom


@needs2
@pytest.mark.parametrize("mode", [1, 3])
@pytest.mark.parametrize("ids_kind", ["identity", "bootstrap"])
def test_devices_option_spreads_one_call_over_two_gpus(lib, oracle, mode, ids_kind):
    spec, chrom = _problem()
    if mode == 3:
        spec = synth.SynthSpec(n_chrom=3, n_inds=7, n_snps=3500, K=6, nsamples=4, rate=4e-7, seed=98, name="multi3")
        chrom = []
        for h in range(spec.n_chrom):
            pos, rate, labels = synth.make_chromosome(spec, h)
            chrom.append({"pos": pos, "rate": rate, "labels": labels})
    K, S, G = spec.K, 3, 110
    dl = spec.donor_label_vec()
    tw = np.random.default_rng(1).random(K * K * S * S)
    ids = ind_id_matrix(spec.n_inds, spec.n_chrom) if ids_kind == "identity" else bootstrap_ids(spec.n_inds, spec.n_chrom, 4)
    oracle.srand(17)
    m_o, raw_o, pairs_o = oracle.data_read_packed(mode, 2, spec.nsamples, chrom, ids, 0.1, G, 10, K, dl, 0.0, 1.0, S, tw, want_raw=True)
    lib.set_option("rand_libc", 0)
    lib.set_option("devices", 2)
    try:
        lib.srand(17)
        m_g, raw_g, st = lib.data_read_packed(mode, 2, spec.nsamples, chrom, ids, 0.1, G, 10, K, dl, 0.0, 1.0, S, tw, want_raw=(mode == 1))
        nxt = lib.rand_fill(3)
    finally:
        lib.set_option("devices", 1)
    assert st["pairs"] == pairs_o
    if mode == 1:
        assert np.array_equal(raw_g, raw_o), int((raw_g != raw_o).sum())
    ok = np.isfinite(m_o) & (m_o != 0)
    assert np.array_equal(np.isfinite(m_g), np.isfinite(m_o))
    assert np.max(np.abs(m_g[ok] - m_o[ok]) / np.abs(m_o[ok])) < 1e-12
    assert np.array_equal(nxt, np.array([oracle.rand() for _ in range(3)], dtype=np.int32))


def _rank_main(rank, world, uid_q, res_q, ids, seed):
    os.environ["CUDA_VISIBLE_DEVICES"] = str(rank)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from fastglobetrotter_b200.companion import PackedCompanion
    lib = PackedCompanion()
    lib.set_option("rand_libc", 0)
    if rank == 0:
        uid = lib.dist_unique_id()
        for _ in range(world - 1):
            uid_q.put(uid)
    else:
        uid = uid_q.get(timeout=120)
    lib.dist_init(rank, world, uid)
    spec, chrom = _problem()
    N, Cn = spec.n_inds, spec.n_chrom
    base, rem = divmod(N, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    K, S, G = spec.K, 3, 110
    dl = spec.donor_label_vec()
    tw = np.random.default_rng(1).random(K * K * S * S)
    my_ids = np.ascontiguousarray(ids.reshape(N, Cn)[lo:hi]).ravel()
    lib.srand(seed)
    m, raw, st = lib.data_read_packed(1, 2, spec.nsamples, chrom, my_ids, 0.1, G, 10, K, dl, 0.0, 1.0, S, tw, want_raw=True,
                                      num_inds_total=N)
    nxt = lib.rand_fill(3)
    # the NULL tensor over the same process group: every rank passes the whole problem, folds its rows, gets the all-reduced tensor
    init = np.random.default_rng(5).random(K * K * G)
    nul, nst = lib.data_read_null_packed(2, spec.nsamples, chrom, np.arange(N), 0.1, G, 10, K, dl, 0.0, 1.0, init=init)
    lib.dist_finalize()
    res_q.put((rank, m, raw, st["pairs"], nxt, nul, nst["pairs"]))


@needs2
def test_one_process_per_gpu_with_the_library_process_group(oracle):
    spec, chrom = _problem()
    N, Cn = spec.n_inds, spec.n_chrom
    K, S, G = spec.K, 3, 110
    dl = spec.donor_label_vec()
    tw = np.random.default_rng(1).random(K * K * S * S)
    ids = bootstrap_ids(N, Cn, 8)
    oracle.srand(23)
    m_o, raw_o, pairs_o = oracle.data_read_packed(1, 2, spec.nsamples, chrom, ids, 0.1, G, 10, K, dl, 0.0, 1.0, S, tw, want_raw=True)
    want_next = np.array([oracle.rand() for _ in range(3)], dtype=np.int32)
    ctx = mp.get_context("spawn")
    uid_q, res_q = ctx.Queue(), ctx.Queue()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, uid_q, res_q, ids, 23)) for r in range(2)]
    for p in procs:
        p.start()
    out = {}
    for _ in range(2):
        r = res_q.get(timeout=300)
        out[r[0]] = r
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n_o, ev = oracle.data_read_null_packed(2, spec.nsamples, chrom, np.arange(N), 0.1, G, 10, K, dl, 0.0, 1.0,
                                           init=np.random.default_rng(5).random(K * K * G))
    for r in (0, 1):
        assert out[r][6] == ev and np.array_equal(out[r][5], n_o), r          # disjoint supports: the all-reduce is exact
    raw = np.concatenate([out[0][2], out[1][2]], axis=0)
    assert out[0][3] + out[1][3] == pairs_o
    assert np.array_equal(raw, raw_o)
    ok = np.isfinite(m_o) & (m_o != 0)
    for r in (0, 1):
        assert np.max(np.abs(out[r][1][ok] - m_o[ok]) / np.abs(m_o[ok])) < 1e-12
        assert np.array_equal(out[r][4], want_next)          # every rank's generator advanced by the whole call's pairs
    assert np.array_equal(out[0][1], out[1][1])              # the all-reduce leaves the same bits on both ranks


@needs2
def test_null_tensor_over_two_gpus_is_bit_exact(lib, oracle):
    """data_read_null with "devices" = 2: every engine folds the label pairs it owns, the partial tensors have disjoint
    supports and their sum is the single-GPU (= reference) tensor bit for bit, including the add-into-caller semantics."""
    spec, chrom = _problem(n_inds=6, n_chrom=2)
    K, G = spec.K, 110
    dl = spec.donor_label_vec()
    init = np.random.default_rng(3).random(K * K * G)
    n_o, ev = oracle.data_read_null_packed(2, spec.nsamples, chrom, np.arange(6), 0.1, G, 10, K, dl, 0.0, 1.0, init=init)
    lib.set_option("devices", 2)
    try:
        n_g, st = lib.data_read_null_packed(2, spec.nsamples, chrom, np.arange(6), 0.1, G, 10, K, dl, 0.0, 1.0, init=init)
    finally:
        lib.set_option("devices", 1)
    assert st["pairs"] == ev
    assert np.array_equal(n_g, n_o)
