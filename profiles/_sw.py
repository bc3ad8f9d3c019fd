import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from fastglobetrotter_b200 import synth, build as fbuild
from fastglobetrotter_b200.companion import PackedCompanion
from fastglobetrotter_b200.synth_torch import make_chromosome_runs_device
lib = PackedCompanion(fbuild.build())
dev = torch.device("cuda", 0)
N, L, K, G = 100, 100_000, 30, 500
spec = synth.SynthSpec(n_chrom=1, n_inds=N, n_snps=L, K=K, nsamples=10, seed=20261021)
pos, rate, run_off, runs = make_chromosome_runs_device(spec, 0, dev)
chrom = [{"pos": pos, "rate": rate, "runs": (run_off, int(runs.data_ptr()))}]
dl = spec.donor_label_vec(); rows = np.arange(N, dtype=np.int32)
ref = None
for nc, ec, cc in ((1,8,16),(4,8,12),(4,8,8),(4,8,6),(1,8,16)):
    lib.set_option("null_consumers", nc); lib.set_option("null_eval_ctas", ec); lib.set_option("null_commit_ctas", cc)
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out, st = lib.data_read_null_packed(2, 10, chrom, rows, 0.1, G, 10, K, dl, 0.0, 1.0)
        dt = time.perf_counter() - t0
    if ref is None: ref = out
    print(f"consumers {nc} eval {ec} commit {cc}: {dt*1e3:.1f} ms  same={np.array_equal(out, ref)}", flush=True)
