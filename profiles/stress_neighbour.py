import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from fastglobetrotter_b200 import synth, build as fbuild
from fastglobetrotter_b200.companion import PackedCompanion
from fastglobetrotter_b200.synth_torch import make_chromosome_device
lib = PackedCompanion(fbuild.build()); lib.set_option("rand_libc", 0)
dev = torch.device("cuda", 0)
N, L, K, G, S, nchr = 100, 100_000, 30, 500, 10, 1
spec = synth.SynthSpec(n_chrom=nchr, n_inds=N, n_snps=L, K=K, nsamples=10, seed=77)
chrom, keep = [], []
for h in range(nchr):
    pos, rate, lab = make_chromosome_device(spec, h, dev); keep.append(lab)
    chrom.append({"pos": pos, "rate": rate, "labels": (lab.data_ptr(), 2, lab.shape[0])})
torch.cuda.synchronize()
dl = spec.donor_label_vec()
tw = np.random.default_rng(0).random(K * K * S * S)
ids = np.arange(N).astype(np.int32)
bp = float(os.environ.get("DIAG_BP", "0"))
lib.set_option("batch_pairs", bp)
lib.set_option("accum_impl", int(os.environ.get("DIAG_IMPL", "2")))
def run():
    lib.srand(11)
    m, raw, st = lib.data_read_packed(1, 2, 10, chrom, ids, 0.1, G, 10, K, dl, 0.0, 1.0, S, tw, want_raw=True)
    return raw.copy(), st
quiet, st = run()
print("launches", st["accum_launches"], "ms", st["ms_total"])
x = torch.empty(1 << 28, dtype=torch.float32, device=dev); y = torch.empty_like(x)
side = torch.cuda.Stream()
for mode in ["copy", "copy", "randn"]:
    for rep in range(3):
        with torch.cuda.stream(side):
            for i in range(40):
                if mode == "copy": y.copy_(x)
                elif mode == "fill": y.fill_(float(i))
                else: y.normal_()
        noisy, st = run()
        torch.cuda.synchronize()
        bad = np.where((noisy != quiet).reshape(N, -1).any(axis=1))[0]
        print(mode, rep, "ms", round(st["ms_total"], 2), "bad individuals:", list(map(int, bad)))
