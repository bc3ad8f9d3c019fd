"""Host-side breakdown of one sharded data_read step (run under torchrun, 2+ ranks): where the weak-scaling overhead goes."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch, torch.distributed as dist
from fastglobetrotter_b200 import synth, shard, build as fbuild
from fastglobetrotter_b200.companion import PackedCompanion
from fastglobetrotter_b200.synth_torch import make_chromosome_device
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
lib = PackedCompanion(fbuild.build()); lib.set_option("device", local); lib.set_option("rand_libc", 0); lib.srand(1)
N, L, K, G, S = 100, 100_000, 30, 500, 10
spec = synth.SynthSpec(n_chrom=1, n_inds=N, n_snps=L, K=K, nsamples=10, seed=20261021)
pos, rate, lab = make_chromosome_device(spec, 0, dev, ind_offset=rank * N)
chrom = [{"pos": pos, "rate": rate, "labels": (lab.data_ptr(), 2, lab.shape[0])}]
dl = spec.donor_label_vec(); tw = np.random.default_rng(0).random(K * K * S * S); ids = np.arange(N).astype(np.int32)
cnt_dev = torch.empty(N, dtype=torch.int64, device=dev); all_dev = torch.empty(world * N, dtype=torch.int64, device=dev)
m_dev = torch.empty(S * S * G, dtype=torch.float64, device=dev)
T = np.zeros(5)
for it in range(13):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    cnt = lib.count_pairs_packed(1, 2, 10, chrom, ids, 0.1, G, 10, K, dl, 0.0, 1.0)
    t1 = time.perf_counter()
    cnt_dev.copy_(torch.from_numpy(cnt.ravel())); dist.all_gather_into_tensor(all_dev, cnt_dev)
    by_rank = list(all_dev.cpu().numpy().reshape(world, 1, N))
    offs, total = shard.stream_offsets(by_rank)
    t2 = time.perf_counter()
    lib.set_option("reuse_prep", 1)
    m, _, st = lib.data_read_packed(1, 2, 10, chrom, ids, 0.1, G, 10, K, dl, 0.0, 1.0, S, tw, num_inds_total=world * N,
                                    pair_offset=offs[rank], pairs_total=total)
    t3 = time.perf_counter()
    m_dev.copy_(torch.from_numpy(m.ravel())); dist.all_reduce(m_dev); out = m_dev.cpu()
    t4 = time.perf_counter()
    if it >= 3: T += [t1 - t0, t2 - t1, t3 - t2, t4 - t3, st["ms_total"] * 1e-3]
if rank == 0:
    T /= 10
    print(f"count pass {T[0]*1e3:.3f} ms | gather+offsets {T[1]*1e3:.3f} | data_read {T[2]*1e3:.3f} (device {T[4]*1e3:.3f}) | all_reduce {T[3]*1e3:.3f} | sum {T[:4].sum()*1e3:.3f}")
dist.destroy_process_group()
