import sys, os, time
sys.path.insert(0, os.getcwd())
import bench, numpy as np, torch
from fastglobetrotter_b200 import synth, build as fbuild
from fastglobetrotter_b200.companion import PackedCompanion
from fastglobetrotter_b200.synth_torch import make_chromosome_device
W = bench.WORK
lib = PackedCompanion(fbuild.build()); lib.set_option("rand_libc", 0); lib.srand(1)
spec = bench.spec_for(W["n_inds"]); dev = torch.device("cuda", 0)
pos, rate, lab = make_chromosome_device(spec, 0, dev)
hl = torch.empty(lab.shape, dtype=torch.int16, pin_memory=True); hl.copy_(lab)
torch.cuda.synchronize()
K, S, G = W["K"], W["S"], W["G"]
tw = lib.temp_weights_for_data_read(K, S, synth.random_weights(K, S))
ids = np.arange(W["n_inds"], dtype=np.int32)
ch_h = [{"pos": pos, "rate": rate, "labels": hl.numpy().view(np.uint16)}]
ch_d = [{"pos": pos, "rate": rate, "labels": (lab.data_ptr(), 2, lab.shape[0])}]
def run(chrom, n=5):
    for _ in range(2): lib.data_read_packed(1, 2, 10, chrom, ids, W["bw"], G, W["grid_start"], K, spec.donor_label_vec(), 0.0, 1.0, S, tw)
    t0 = time.perf_counter()
    for _ in range(n): m, _, st = lib.data_read_packed(1, 2, 10, chrom, ids, W["bw"], G, W["grid_start"], K, spec.donor_label_vec(), 0.0, 1.0, S, tw)
    return (time.perf_counter() - t0) / n * 1e3, st, m
for split in (1, 2, 3, 4, 6):
    lib.set_option("split_ready", split)
    ms, st, m = run(ch_d)
    print(f"device split_ready={split}: {ms:.2f} ms launches={st['accum_launches']} accum={st['ms_accum']:.2f} rand={st['ms_rand']:.2f} proj={st['ms_project']:.2f} chunk={st['ms_chunk']:.2f}")
lib.set_option("split_ready", 1)
for ranges in (2, 4, 8):
    lib.set_option("ranges", ranges)
    ms, st, m = run(ch_h)
    print(f"host ranges={ranges}: {ms:.2f} ms launches={st['accum_launches']}")
