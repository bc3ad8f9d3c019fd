import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from fastglobetrotter_b200 import synth, build as fbuild
from fastglobetrotter_b200.companion import PackedCompanion
from fastglobetrotter_b200.synth_torch import make_chromosome_device
lib = PackedCompanion(fbuild.build()); lib.set_option("rand_libc", 0)
dev = torch.device("cuda", 0)
def case(N, L, K, G, S, nchr, boot, tag):
    spec = synth.SynthSpec(n_chrom=nchr, n_inds=N, n_snps=L, K=K, nsamples=10, seed=77)
    chrom, keep = [], []
    for h in range(nchr):
        pos, rate, lab = make_chromosome_device(spec, h, dev); keep.append(lab)
        chrom.append({"pos": pos, "rate": rate, "labels": (lab.data_ptr(), 2, lab.shape[0])})
    torch.cuda.synchronize()
    dl = spec.donor_label_vec()
    tw = np.random.default_rng(0).random(K * K * S * S)
    rng = np.random.default_rng(5)
    ids = (np.sort(rng.integers(0, N, size=(nchr, N)), axis=1).T.ravel() if boot else np.repeat(np.arange(N), nchr)).astype(np.int32)
    outs = []
    for rep in range(3):
        lib.srand(11)
        m, _, st = lib.data_read_packed(1, 2, 10, chrom, ids, 0.1, G, 10, K, dl, 0.0, 1.0, S, tw)
        outs.append(m)
    lib.set_option("accum_impl", 1)
    lib.srand(11)
    m1, _, st1 = lib.data_read_packed(1, 2, 10, chrom, ids, 0.1, G, 10, K, dl, 0.0, 1.0, S, tw)
    lib.set_option("accum_impl", 2)
    print(f"{tag}: N={N} K={K} chr={nchr} boot={boot} pairs={st['pairs']:.2e} launches={st['accum_launches']} "
          f"rep01={np.array_equal(outs[0], outs[1])} rep12={np.array_equal(outs[1], outs[2])} vs_impl1={np.array_equal(outs[0], m1)} "
          f"maxrel={np.nanmax(np.abs(outs[0]-m1)/np.abs(m1)):.2e}")
pass
pass
pass
pass
pass
if os.environ.get("DIAG_SMALL"):
    lib.set_option("batch_pairs", float(os.environ.get("DIAG_BP", "4e6")))
    case(24, 30_000, 12, 300, 4, 2, False, "small 2chr batched")
    sys.exit(0)
print("---- forced batches")
lib.set_option("batch_pairs", 1.2e8)
case(300, 100_000, 30, 500, 10, 1, False, "N300 batched")
case(100, 100_000, 30, 500, 10, 2, False, "2chr batched")
lib.set_option("batch_pairs", 0)
