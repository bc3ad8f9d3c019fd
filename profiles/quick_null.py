import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from fastglobetrotter_b200 import synth, build as fbuild
from fastglobetrotter_b200.companion import PackedCompanion
from fastglobetrotter_b200.synth_torch import make_chromosome_device
lib = PackedCompanion(fbuild.build())
dev = torch.device("cuda", 0)
N, L, K, G, nchr = 100, 7000, 26, 291, (int(sys.argv[1]) if len(sys.argv) > 1 else 2)
spec = synth.SynthSpec(n_chrom=nchr, n_inds=N, n_snps=L, K=K, nsamples=1, rate=60.0/L/1000/100, seed=5)
chrom = []; keep = []
for h in range(nchr):
    pos, rate, lab = make_chromosome_device(spec, h, dev); keep.append(lab)
    chrom.append({"pos": pos, "rate": rate, "labels": (lab.data_ptr(), 2, lab.shape[0])})
torch.cuda.synchronize()
dl = spec.donor_label_vec()
for it in range(2):
    t0 = time.perf_counter()
    out, st = lib.data_read_null_packed(2, 1, chrom, np.arange(N), 0.1, G, 10, K, dl, 0.0, 1.0)
    dt = time.perf_counter() - t0
print(f"null N={N} L={L} K={K} G={G} chr={nchr}: {dt*1e3:.1f} ms  evals={st['pairs']:.3e} accepted={st['accepted']:.3e} chunks={st['chunks']}")
lab_h = keep[0].cpu().numpy().view(np.uint16)
pops = dl[lab_h[:, ::50].ravel() - 1]
print("label freq top5:", np.sort(np.bincount(pops, minlength=K+1) / pops.size)[::-1][:5])
