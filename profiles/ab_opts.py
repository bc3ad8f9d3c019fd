"""A/B of engine options on the bench workload (RLE paintings resident in HBM): per option set, the accumulation time
(CUDA events inside the library), the call time and a hash of the curves.  OPTSETS="a=1,b=2;c=3" (';' separates sets)."""
import sys, os, hashlib, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from fastglobetrotter_b200 import synth, build as fbuild
from fastglobetrotter_b200.companion import PackedCompanion
from fastglobetrotter_b200.synth_torch import make_chromosome_runs_device
lib = PackedCompanion(fbuild.build()); lib.set_option("rand_libc", 0)
dev = torch.device("cuda", 0)
N, L, K, G, S = [int(os.environ.get(k, v)) for k, v in (('N', 100), ('L', 100_000), ('K', 30), ('G', 500), ('S', 10))]
MODE = int(os.environ.get('MODE', 1))
spec = synth.SynthSpec(n_chrom=1, n_inds=N, n_snps=L, K=K, nsamples=10, seed=20261021)
pos, rate, run_off, runs = make_chromosome_runs_device(spec, 0, dev)
chrom = [{"pos": pos, "rate": rate, "runs": (run_off, int(runs.data_ptr()))}]
if os.environ.get("HOST"):         # the e2e leg of bench.py: the runs in pinned host memory, copied by the library inside the call
    hr = torch.empty(runs.shape, dtype=runs.dtype, pin_memory=True); hr.copy_(runs); torch.cuda.synchronize()
    chrom = [{"pos": pos, "rate": rate, "runs": (run_off, hr.numpy().view(np.uint32).reshape(-1, 2))}]
dl = spec.donor_label_vec(); tw = lib.temp_weights_for_data_read(K, S, synth.random_weights(K, S)); ids = np.arange(N).astype(np.int32)
defaults = {}
for oset in os.environ.get("OPTSETS", "").split(";") or [""]:
    kv = [x.split("=") for x in oset.split(",") if x]
    for k, v in kv:
        defaults.setdefault(k, lib.get_option(k)); lib.set_option(k, float(v))
    acc, tot, prj = [], [], []
    for rep in range(8):
        lib.srand(11); torch.cuda.synchronize(); t0 = time.perf_counter()
        m, _, st = lib.data_read_packed(MODE, 2, 10, chrom, ids, 0.1, G, 10, K, dl, 0.0, 1.0, S, tw)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        if rep >= 3: acc.append(st["ms_accum"]); tot.append((t1 - t0) * 1e3); prj.append(st.get("ms_project", 0.0))
    print(f"{oset or 'default':40s} ms_accum {np.mean(acc):.3f} (min {np.min(acc):.3f})  call {np.mean(tot):.3f} ms  project {np.mean(prj):.3f}  pairs {st['pairs']}  sha {hashlib.sha256(np.ascontiguousarray(m).tobytes()).hexdigest()[:12]}", flush=True)
    for k, v in defaults.items(): lib.set_option(k, v)
