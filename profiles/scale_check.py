"""One-off scale validation (config-3-like, scaled to one GPU): N=500 individuals x 2 chromosomes x 100k SNPs, K=50.
Checks that the batched/pipelined paths reproduce the unbatched result bit-for-bit at a size the oracle cannot reach,
and prints throughput."""
import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from fastglobetrotter_b200 import synth, build as fbuild
from fastglobetrotter_b200.companion import PackedCompanion
from fastglobetrotter_b200.synth_torch import make_chromosome_device
lib = PackedCompanion(fbuild.build()); lib.set_option("rand_libc", 0)
dev = torch.device("cuda", 0)
N, L, K, G, S, nchr = 500, 100_000, 50, 500, 12, 2
spec = synth.SynthSpec(n_chrom=nchr, n_inds=N, n_snps=L, K=K, nsamples=10, seed=77)
chrom, keep = [], []
for h in range(nchr):
    pos, rate, lab = make_chromosome_device(spec, h, dev); keep.append(lab)
    chrom.append({"pos": pos, "rate": rate, "labels": (lab.data_ptr(), 2, lab.shape[0])})
torch.cuda.synchronize()
dl = spec.donor_label_vec()
tw = np.random.default_rng(0).random(K * K * S * S)
rng = np.random.default_rng(5)
ids = np.sort(rng.integers(0, N, size=(nchr, N)), axis=1).T.ravel().astype(np.int32)      # a bootstrap replicate
res = {}
for name, opts in (("default", {}), ("batched", {"batch_pairs": 9e7}), ("mode3", {})):
    for k, v in opts.items(): lib.set_option(k, v)
    mode = 3 if name == "mode3" else 1
    lib.srand(11)
    t0 = time.perf_counter()
    m, _, st = lib.data_read_packed(mode, 2, 10, chrom, ids, 0.1, G, 10, K, dl, 0.0, 1.0, S, tw)
    dt = time.perf_counter() - t0
    lib.srand(11)
    t0 = time.perf_counter()
    m2, _, st = lib.data_read_packed(mode, 2, 10, chrom, ids, 0.1, G, 10, K, dl, 0.0, 1.0, S, tw)
    dt = time.perf_counter() - t0
    for k in opts: lib.set_option(k, 0)
    res[name] = m
    print(f"{name}: {dt*1e3:.1f} ms, pairs={st['pairs']:.3e} -> {st['pairs']/dt/1e9:.1f} G pairs/s, accum launches={st['accum_launches']}, "
          f"repeatable={np.array_equal(m, m2)}, finite={np.isfinite(m).all()}")
print("batched == default:", np.array_equal(res["default"], res["batched"]))
print("free/total GB:", [x / 1e9 for x in torch.cuda.mem_get_info()])
