"""time data_read_null on the bench chromosome for a few task-table settings (OPTS-like sweep of null_slab_bins)"""
import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from fastglobetrotter_b200 import synth, build as fbuild
from fastglobetrotter_b200.companion import PackedCompanion
from fastglobetrotter_b200.synth_torch import make_chromosome_device
lib = PackedCompanion(fbuild.build())
dev = torch.device("cuda", 0)
N, L, K, G = int(os.environ.get("N", 40)), int(os.environ.get("L", 100_000)), int(os.environ.get("K", 30)), int(os.environ.get("G", 500))
spec = synth.SynthSpec(n_chrom=1, n_inds=N, n_snps=L, K=K, nsamples=10, seed=20261021)
pos, rate, lab = make_chromosome_device(spec, 0, dev)
chrom = [{"pos": pos, "rate": rate, "labels": (lab.data_ptr(), 2, lab.shape[0])}]
dl = spec.donor_label_vec(); rows = np.arange(N, dtype=np.int32)
ref = None
for sb in [int(x) for x in os.environ.get("SB", "0,2,4,8,16,32,64").split(",")]:
    lib.set_option(os.environ.get("KEY", "null_slab_bins"), sb)
    lib.data_read_null_packed(2, 10, chrom, rows, 0.1, G, 10, K, dl, 0.0, 1.0)
    t0 = time.perf_counter()
    out, st = lib.data_read_null_packed(2, 10, chrom, rows, 0.1, G, 10, K, dl, 0.0, 1.0)
    dt = time.perf_counter() - t0
    if ref is None: ref = out
    print(f"{os.environ.get('KEY', 'null_slab_bins')}={sb:3d}: {dt*1e3:8.1f} ms (device {st['ms_total']:.1f}, chunking {st['ms_chunk']:.1f}, kernel {st['ms_accum']:.1f})  "
          f"{st['pairs']/dt/1e9:6.1f} G evals/s  same={np.array_equal(out, ref)}")
lib.set_option(os.environ.get("KEY", "null_slab_bins"), 0)
