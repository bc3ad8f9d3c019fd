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
torch.cuda.synchronize()
K, S, G = W["K"], W["S"], W["G"]
tw = lib.temp_weights_for_data_read(K, S, synth.random_weights(K, S))
ids = np.arange(W["n_inds"], dtype=np.int32)
ch_d = [{"pos": pos, "rate": rate, "labels": (lab.data_ptr(), 2, lab.shape[0])}]
def run(chrom, n=8):
    for _ in range(2): lib.data_read_packed(1, 2, 10, chrom, ids, W["bw"], G, W["grid_start"], K, spec.donor_label_vec(), 0.0, 1.0, S, tw)
    t0 = time.perf_counter()
    for _ in range(n): m, _, st = lib.data_read_packed(1, 2, 10, chrom, ids, W["bw"], G, W["grid_start"], K, spec.donor_label_vec(), 0.0, 1.0, S, tw)
    return (time.perf_counter() - t0) / n * 1e3, st, m
for blocks, ranges in ((1, 1), (2, 2), (2, 1), (3, 3), (4, 2), (4, 4)):
    lib.set_option("blocks", blocks); lib.set_option("ranges", ranges)
    ms, st, m = run(ch_d)
    print(f"device blocks={blocks} ranges={ranges}: {ms:.2f} ms launches={st['accum_launches']} accum={st['ms_accum']:.2f} chunk={st['ms_chunk']:.2f}")
