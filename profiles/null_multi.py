"""data_read_null on the bench chromosome (100 recipients) from run-length-encoded paintings in pinned host memory, on 1 and on
DEVICES GPUs of one process ("devices" option: rows of A labels per GPU, partial tensors with disjoint supports)."""
import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from fastglobetrotter_b200 import synth, build as fbuild
from fastglobetrotter_b200.companion import PackedCompanion
from fastglobetrotter_b200.synth_torch import make_chromosome_runs_device
lib = PackedCompanion(fbuild.build())
dev = torch.device("cuda", 0)
N, L, K, G = [int(os.environ.get(k, v)) for k, v in (("N", 100), ("L", 100_000), ("K", 30), ("G", 500))]
spec = synth.SynthSpec(n_chrom=1, n_inds=N, n_snps=L, K=K, nsamples=10, seed=20261021)
pos, rate, run_off, runs = make_chromosome_runs_device(spec, 0, dev)
hr = torch.empty(runs.shape, dtype=runs.dtype, pin_memory=True); hr.copy_(runs); torch.cuda.synchronize()
chrom = [{"pos": pos, "rate": rate, "runs": (run_off, hr.numpy().view(np.uint32).reshape(-1, 2))}]
dl = spec.donor_label_vec(); rows = np.arange(N, dtype=np.int32)
ref = None
for nd in [int(x) for x in os.environ.get("DEVS", "1,2").split(",")]:
    lib.set_option("devices", nd)
    for rep in range(3):
        lib.set_option("verbose", 1 if (rep == 2 and os.environ.get("VERBOSE")) else 0)
        t0 = time.perf_counter()
        out, st = lib.data_read_null_packed(2, 10, chrom, rows, 0.1, G, 10, K, dl, 0.0, 1.0)
        dt = time.perf_counter() - t0
    lib.set_option("verbose", 0)
    if ref is None: ref = out
    print(f"devices={nd}: {dt*1e3:8.1f} ms  {st['pairs']/dt/1e9:6.1f} G evals/s  same bits as one device: {np.array_equal(out, ref)}", flush=True)
lib.set_option("devices", 1)
