"""data_read_null sharded by label-pair ownership (run under torchrun): every rank holds the recipients' paintings, folds
the label pairs it owns, one all-reduce(sum) of the K x K x G tensor (disjoint supports -> exact).  Prints the time of
the sharded call and checks the result against an unsharded call on rank 0."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch, torch.distributed as dist
from fastglobetrotter_b200 import synth, build as fbuild
from fastglobetrotter_b200.companion import PackedCompanion
from fastglobetrotter_b200.synth_torch import make_chromosome_device
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
lib = PackedCompanion(fbuild.build()); lib.set_option("device", local)
N, L, K, G = int(os.environ.get("N", 60)), 100_000, 30, 500
spec = synth.SynthSpec(n_chrom=1, n_inds=N, n_snps=L, K=K, nsamples=10, seed=20261021)
pos, rate, lab = make_chromosome_device(spec, 0, dev)
chrom = [{"pos": pos, "rate": rate, "labels": (lab.data_ptr(), 2, lab.shape[0])}]
dl = spec.donor_label_vec(); rows = np.arange(N, dtype=np.int32)
def run(w, r):
    lib.set_option("null_world", w); lib.set_option("null_rank", r)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    part, st = lib.data_read_null_packed(2, 10, chrom, rows, 0.1, G, 10, K, dl, 0.0, 1.0)
    t = torch.from_numpy(part).to(dev)
    if w > 1: dist.all_reduce(t)
    out = t.cpu().numpy()
    torch.cuda.synchronize(); dist.barrier()
    return out, time.perf_counter() - t0, st
run(world, rank)
sharded, dt_s, st = run(world, rank)
full, dt_f, _ = run(1, 0)
if rank == 0:
    print(f"{world} ranks: sharded {dt_s*1e3:.1f} ms vs one GPU {dt_f*1e3:.1f} ms ({st['pairs']:.3e} chunk-pair evaluations); "
          f"bit-identical: {np.array_equal(sharded, full)}")
lib.set_option("null_world", 1); lib.set_option("null_rank", 0)
dist.destroy_process_group()
