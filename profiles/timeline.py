"""Print the engine's verbose host/device timeline for one data_read on the bench workload (device-resident paintings;
RLE=1: run-length-encoded rows, the bench's default input)."""
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from fastglobetrotter_b200 import synth, build as fbuild
from fastglobetrotter_b200.companion import PackedCompanion
from fastglobetrotter_b200.synth_torch import make_chromosome_device, make_chromosome_runs_device
lib = PackedCompanion(fbuild.build()); lib.set_option("rand_libc", 0)
dev = torch.device("cuda", 0)
N, L, K, G, S = 100, 100_000, 30, 500, 10
N, L, K, G, S = [int(os.environ.get(k, v)) for k, v in (('N', N), ('L', L), ('K', K), ('G', G), ('S', S))]
MODE = int(os.environ.get('MODE', 1))
spec = synth.SynthSpec(n_chrom=1, n_inds=N, n_snps=L, K=K, nsamples=10, seed=20261021)
pos, rate, lab = make_chromosome_device(spec, 0, dev)
chrom = [{"pos": pos, "rate": rate, "labels": (lab.data_ptr(), 2, lab.shape[0])}]
if os.environ.get("RLE"):
    pos, rate, run_off, runs = make_chromosome_runs_device(spec, 0, dev)
    chrom = [{"pos": pos, "rate": rate, "runs": (run_off, int(runs.data_ptr()))}]
elif os.environ.get("HOST"):
    hl = torch.empty(lab.shape, dtype=torch.int16, pin_memory=True); hl.copy_(lab)
    chrom = [{"pos": pos, "rate": rate, "labels": hl.numpy().view(np.uint16)}]
for k, v in [kv.split("=") for kv in os.environ.get("OPTS", "").split(",") if kv]:
    lib.set_option(k, float(v))
torch.cuda.synchronize()
dl = spec.donor_label_vec(); tw = lib.temp_weights_for_data_read(K, S, synth.random_weights(K, S)); ids = np.arange(N).astype(np.int32)
for rep in range(4):
    lib.set_option("verbose", 1 if rep == 3 else 0)
    lib.srand(11)
    m, _, st = lib.data_read_packed(MODE, 2, 10, chrom, ids, 0.1, G, 10, K, dl, 0.0, 1.0, S, tw)
print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items()})
