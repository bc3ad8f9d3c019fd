#!/usr/bin/env python3
"""bench.py -- coancestry-curve hot path throughput on B200 (see DESIGN.md "Measurement").

One "step" = one full `data_read` pass (chunk paintings -> sampling schedule -> glibc rand stream ->
order-preserving fp64 accumulation -> projection -> normalised curves) over the synthetic workload
of BASELINE.json configs[1]: 1 chromosome, 200 target haplotypes (100 diploid individuals) x 10
painting samples, 100k SNPs, K = 30 donor groups, 500 bins of 0.1 cM, per GPU (weak scaling).

  value : sampled chunk-pair updates per second, inputs (decoded paintings) resident in HBM
  e2e   : the same through the packed C ABI with HOST (pinned) paintings: H2D of the labels and D2H
          of the curves inside the timed region
  --impl reference : the compiled reference companion (oracle/_ref) on host cores, bounded sample
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "coancestry_chunk_pair_updates_per_sec"
UNIT = "pair-updates/s"

WORK = dict(n_chrom=1, n_inds=100, ploidy=2, nsamples=10, n_snps=100_000, K=30, G=500, bw=0.1, grid_start=10, S=10,
            haps_per_pop=20, weight_min=0.0, cutoff=1.0, mode=1)
WORKLOAD_NAME = ("configs[1]: synthetic 1 chromosome, 200 target haps (100 diploid inds) x 10 samples, 100k SNPs, "
                 "K=30 donor groups, 500 x 0.1 cM bins, S=10 surrogates, data_read (mode 1) semantics, per GPU")


def spec_for(n_inds, seed=20261019 + 2):
    from fastglobetrotter_b200 import synth
    return synth.SynthSpec(n_chrom=WORK["n_chrom"], n_inds=n_inds, ploidy=WORK["ploidy"], nsamples=WORK["nsamples"],
                           n_snps=WORK["n_snps"], K=WORK["K"], haps_per_pop=WORK["haps_per_pop"], seed=seed, name="cfg2")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML polled every few milliseconds from a thread
    (the timed region of the default run is ~60 ms, too short for nvidia-smi's own loop), nvidia-smi -lms as fallback."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
            0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.nvml, self.handle, self.samples, self.mask, self.stop_flag = None, None, [], 0, False

    def _nvml_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.index < len(ids) and ids[self.index].isdigit():
                return int(ids[self.index])
        return self.index

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._nvml_index())
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                          str(self.index), "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                                         text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                self.samples.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                self.mask |= int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
            except Exception:
                pass
            time.sleep(0.003)

    def _pump(self):
        for ln in self.proc.stdout:
            self.rows.append(ln.strip())

    def stop(self):
        if self.nvml:
            self.stop_flag = True
            self.t.join(timeout=1)
            reasons = sorted(nm for bit, nm in self.BITS.items() if self.mask & bit)
            return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                    "reasons": reasons, "samples": len(self.samples), "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvidia-smi"}


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from fastglobetrotter_b200 import build as fbuild
    from fastglobetrotter_b200 import synth
    from fastglobetrotter_b200.companion import PackedCompanion
    from fastglobetrotter_b200.synth_torch import make_chromosome_device, make_chromosome_runs_device

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = PackedCompanion(fbuild.build())
    lib.set_option("device", local)
    lib.set_option("rand_libc", 0)
    lib.srand(1)
    for kv in filter(None, os.environ.get("FGT_OPTS", "").split(",")):      # tuning / A-B runs: FGT_OPTS=key=value,...
        k, v = kv.split("=")
        lib.set_option(k.strip(), float(v))

    W = WORK
    spec = spec_for(W["n_inds"])
    K, S, G = W["K"], W["S"], W["G"]
    # paintings in the library's packed format: run-length-encoded rows (fgt_run_t), generated on the device
    chrom_dev, chrom_host, keep = [], [], []
    n_runs_total = 0
    for h in range(W["n_chrom"]):
        pos, rate, run_off, runs = make_chromosome_runs_device(spec, h, dev, ind_offset=rank * W["n_inds"])
        keep.append(runs)
        n_runs_total += int(run_off[-1])
        chrom_dev.append({"pos": pos, "rate": rate, "runs": (run_off, int(runs.data_ptr()))})
        hr = torch.empty(runs.shape, dtype=torch.int32, pin_memory=True)
        hr.copy_(runs)
        keep.append(hr)
        chrom_host.append({"pos": pos, "rate": rate, "runs": (run_off, hr.numpy().view(np.uint32).reshape(-1, 2))})
    torch.cuda.synchronize()
    dl = spec.donor_label_vec()
    weights = synth.random_weights(K, S)
    tw = lib.temp_weights_for_data_read(K, S, weights)
    N = W["n_inds"]
    ids = np.repeat(np.arange(N, dtype=np.int32), W["n_chrom"])
    common = (W["mode"], W["ploidy"], W["nsamples"])
    tail = (ids, W["bw"], G, W["grid_start"], K, dl, W["weight_min"], W["cutoff"])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    if world > 1:
        # the library's own process group: rank 0 draws an NCCL id, torch.distributed only carries its 128 bytes
        uid = [lib.dist_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        lib.dist_init(rank, world, uid[0])

    def one_step(chrom):
        """one data_read pass of this rank's shard; returns (curves, stats).  With several ranks the call is collective:
        pair counts are all-gathered and turned into stream offsets on the device, the partial curves all-reduced,
        all inside the library (include/fgt_companion.h, fgt_dist_init)."""
        m, _, st = lib.data_read_packed(*common, chrom, *tail, S, tw, num_inds_total=world * N)
        return m, st

    def timed(chrom, steps, warmup):
        lib.srand(1)                             # every timed series draws the same stream: results are comparable bit for bit
        for _ in range(warmup):
            one_step(chrom)
        barrier()
        t0 = time.perf_counter()
        agg = {"pairs": 0, "accepted": 0, "ms_accum": 0.0, "accum_launches": 0, "kernel_launches": 0, "chunks": 0,
               "h2d_bytes": 0, "d2h_bytes": 0, "ms_total": 0.0, "ms_rand": 0.0, "ms_chunk": 0.0, "ms_project": 0.0}
        for _ in range(steps):
            m, st = one_step(chrom)
            for k in agg:
                agg[k] += st[k]
        barrier()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt, float(agg["pairs"])], dtype=torch.float64, device=dev)
        if world > 1:
            tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
            dt, total_pairs = float(tmax[0]), float(tsum[1])
        else:
            total_pairs = float(agg["pairs"])
        return dt, total_pairs, agg, m

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    dt, total_pairs, agg, means = timed(chrom_dev, args.steps, args.warmup)
    clocks = sampler.stop() if sampler else None
    e_steps = max(1, min(args.steps, 5))
    dt_e, pairs_e, agg_e, _ = timed(chrom_host, e_steps, 2)
    # relaxed mode (opt-in, fp64 atomics): the regime SURVEY 8(d)'s 16 B/update roofline describes; reported beside strict
    relaxed = None
    if world == 1:
        lib.set_option("strict", 0)
        dt_r, pairs_r, agg_r, means_r = timed(chrom_dev, e_steps, 2)
        lib.set_option("strict", 1)
        # same stream position as a strict call: one call each right after srand(1)
        lib.srand(1)
        m_s, _ = one_step(chrom_dev)
        lib.set_option("strict", 0)
        lib.srand(1)
        m_r, _ = one_step(chrom_dev)
        lib.set_option("strict", 1)
        ok = np.isfinite(m_s) & (m_s != 0)
        relaxed = {"dt": dt_r, "pairs": pairs_r, "agg": agg_r, "steps": e_steps,
                   "max_rel_diff_vs_strict": float(np.max(np.abs(m_r[ok] - m_s[ok]) / np.abs(m_s[ok])))}
    variants = None
    if world == 1 and args.variants:
        variants = {}
        for name, opts in (("bulk_rand", {"rand_impl": 0}), ("lane_queue_commit", {"commit_impl": 1})):
            for k, v in opts.items():
                lib.set_option(k, v)
            dt_v, pairs_v, agg_v, means_v = timed(chrom_dev, e_steps, 2)
            for k in opts:
                lib.set_option(k, {"rand_impl": 1, "commit_impl": 0}[k])
            variants[name] = {"value": pairs_v / dt_v, "ms_per_step": dt_v / e_steps * 1e3, "ms_accum": agg_v["ms_accum"] / e_steps,
                              "ms_rand": agg_v["ms_rand"] / e_steps}
    dense = None
    if world == 1 and not args.no_dense:
        # the same paintings as dense uint16 label matrices (round 1's packed format): chunked on the device by the label-scan kernels
        dchrom_dev, dchrom_host = [], []
        for h in range(W["n_chrom"]):
            pos, rate, lab = make_chromosome_device(spec, h, dev, ind_offset=rank * W["n_inds"])
            keep.append(lab)
            dchrom_dev.append({"pos": pos, "rate": rate, "labels": (lab.data_ptr(), 2, lab.shape[0])})
            hl = torch.empty(lab.shape, dtype=torch.int16, pin_memory=True)
            hl.copy_(lab)
            keep.append(hl)
            dchrom_host.append({"pos": pos, "rate": rate, "labels": hl.numpy().view(np.uint16)})
        lib.encode = "dense"
        dt_d, pairs_d, agg_d, means_d = timed(dchrom_dev, 3, 2)
        dt_dh, pairs_dh, agg_dh, _ = timed(dchrom_host, 3, 1)
        lib.encode = "rle"
        dense = {"value": pairs_d / dt_d, "ms_per_step": dt_d / 3 * 1e3, "e2e_value": pairs_dh / dt_dh, "e2e_ms_per_step": dt_dh / 3 * 1e3,
                 "e2e_h2d_bytes_per_step": agg_dh["h2d_bytes"] / 3}

    null_dist = None
    if world > 1 and not os.environ.get("FGT_BENCH_NO_NULL"):
        # ONE data_read_null job (the first 100 recipients of rank 0's paintings) over all ranks: every rank passes the whole problem,
        # folds its rows of the K x K x G tensor, one ncclAllReduce inside the library -- strong scaling of the NULL path
        nchrom = []
        for h in range(W["n_chrom"]):
            pos, rate, run_off, runs = make_chromosome_runs_device(spec, h, dev, ind_offset=0)
            keep.append(runs)
            nchrom.append({"pos": pos, "rate": rate, "runs": (run_off, int(runs.data_ptr()))})
        rows = np.arange(min(100, N), dtype=np.int32)
        nargs = (W["ploidy"], W["nsamples"], nchrom, rows, W["bw"], G, W["grid_start"], K, dl, W["weight_min"], W["cutoff"])
        lib.data_read_null_packed(*nargs)
        barrier()
        t0 = time.perf_counter()
        nul, nst = lib.data_read_null_packed(*nargs)
        barrier()
        dt_n = time.perf_counter() - t0
        tn = torch.tensor([dt_n], dtype=torch.float64, device=dev)
        dist.all_reduce(tn, op=dist.ReduceOp.MAX)
        null_dist = {"value": nst["pairs"] / float(tn[0]), "unit": "chunk-pair evaluations/s", "ms": float(tn[0]) * 1e3, "n_gpus": world,
                     "scaling": "strong", "individuals": int(rows.size), "pair_evaluations": nst["pairs"], "checksum": float(np.sum(nul))}

    out = None
    if rank == 0:
        peak, peak_src = peaks()
        P = K * (K + 1) // 2
        launches = max(1, agg["accum_launches"])
        per_launch_pairs = agg["pairs"] / launches
        per_launch_chunks = agg["chunks"] / launches
        alg_bytes = 16.0 * per_launch_pairs + 32.0 * per_launch_chunks + 8.0 * P * G * N
        kern_s = (agg["ms_accum"] / launches) * 1e-3
        achieved = alg_bytes / kern_s / 1e9 if kern_s > 0 else 0.0
        traffic = None
        tp = os.path.join(ROOT, "profiles", "accum_dram_bytes.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        out = {
            "metric": METRIC, "value": total_pairs / dt, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic (device-generated paintings, seed 20261021)",
            "config": {"workload": WORKLOAD_NAME, "pairs_per_step_per_gpu": agg["pairs"] / args.steps,
                       "paintings": "run-length encoded rows (the parse layer's output), resident in HBM for `value`, pinned host memory for `e2e`",
                       "accepted_fraction": agg["accepted"] / max(1, agg["pairs"]),
                       "accumulation": "strict (order-preserving fp64 fold, bit-exact vs reference)",
                       "l2": "working set > L2: every step writes and re-reads a fresh 2.4 GB update-record stream (126 MB L2) and a 186 MB tensor"},
            "e2e": {"value": pairs_e / dt_e, "unit": UNIT, "h2d_bytes_per_step": agg_e["h2d_bytes"] / e_steps,
                    "d2h_bytes_per_step": agg_e["d2h_bytes"] / e_steps, "steps": e_steps,
                    "ms_per_step": dt_e / e_steps * 1e3, "api": "fgt_data_read_packed (run-length-encoded paintings in pinned host memory)"},
            "gpu_launches": int(agg["kernel_launches"]),
            "roofline": {"bound": "hbm", "kernel": "k_eval_pairs + k_commit_ordered (strict accumulation, phase 1 + phase 2)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": kern_s * 1e3},
            "paintings": {"format": "run-length encoded (fgt_run_t, 8 B per run)", "runs_per_gpu": n_runs_total,
                          "bytes_per_gpu": 8 * n_runs_total},
            "device_ms_per_step": {k: agg[k] / args.steps for k in ("ms_total", "ms_chunk", "ms_rand", "ms_accum", "ms_project")},
            "clocks": clocks,
            "means_checksum": float(np.nansum(means)),
        }
        if relaxed is not None:
            la = max(1, relaxed["agg"]["accum_launches"])
            ks = relaxed["agg"]["ms_accum"] / la * 1e-3
            ach = alg_bytes / ks / 1e9 if ks > 0 else 0.0
            tr = None
            if os.path.exists(tp):
                try:
                    tr = json.load(open(tp)).get("relaxed_dram_bytes_per_launch")
                except Exception:
                    tr = None
            out["roofline_relaxed"] = {"bound": "hbm", "kernel": "k_eval_pairs<relaxed> (opt-in strict=0: red.global.add.f64 into the tensor, not bit-exact)",
                                       "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": tr,
                                       "kernel_ms": ks * 1e3, "value": relaxed["pairs"] / relaxed["dt"],
                                       "ms_per_step": relaxed["dt"] / relaxed["steps"] * 1e3,
                                       "max_rel_diff_of_curves_vs_strict": relaxed["max_rel_diff_vs_strict"]}
        if variants is not None:
            out["variants"] = variants
        if dense is not None:
            out["dense_label_input"] = dense
        if world == 1:
            out["other_paths"] = other_paths(lib, chrom_dev, common, tail, S, tw, N, dl, K, G, W)
        elif null_dist is not None:
            out["other_paths"] = {"data_read_null": null_dist}
        if not args.no_cpu_baseline and world == 1:
            with tempfile.TemporaryDirectory() as tmp:
                cb, files0, means_ref = cpu_baseline(sample_inds=args.cpu_sample_inds, procs=1, steps=3, warmup=0, tmp=tmp)
                out["cpu_baseline"] = cb
                if files0 is not None:
                    same, fe = file_parity_and_e2e(lib, files0, means_ref, args.cpu_sample_inds, cb.get("ms_per_call"))
                    out["parity_checked"] = same
                    out["parity"] = ("curves of .C data_read (16 individuals) and the NULL tensor of .C data_read_null (8 individuals) on the cpu_baseline leg's files == compiled reference, bit for bit"
                                     if same else "MISMATCH between the GPU library and the compiled reference on the cpu_baseline files")
                    out["file_e2e"] = fe
                else:
                    out["parity_checked"] = False
    if world > 1:
        dist.barrier()
        lib.dist_finalize()
        dist.destroy_process_group()
    if out is not None:
        print(json.dumps(out))


# ------------------------------------------------------------------------------------------------
def _ref_worker(argtuple):
    """one process: compiled reference data_read on its own sample files.
    Returns (pairs drawn by the timed calls, [seconds per timed call], means of the very first call)."""
    lib_path, sl, rl, n_inds, rec_labels, steps, warmup = argtuple
    import ctypes as C
    sys.path.insert(0, ROOT)
    from fastglobetrotter_b200.companion import Companion
    from fastglobetrotter_b200 import synth
    W = WORK
    ref = Companion(lib_path)
    K, S, G = W["K"], W["S"], W["G"]
    weights = synth.random_weights(K, S)
    tw = ref.temp_weights_for_data_read(K, S, weights)
    ids = np.repeat(np.arange(n_inds, dtype=np.int32), W["n_chrom"])
    dl = spec_for(n_inds).donor_label_vec()
    pairs = None
    try:
        pairs = C.c_long.in_dll(ref.lib, "REF_DUMP_pairs")
    except ValueError:
        pass
    C.CDLL(None).srand(1)
    times, first, p0 = [], None, 0
    for it in range(warmup + steps):
        if it == warmup and pairs is not None:
            p0 = pairs.value
        t0 = time.perf_counter()
        m = ref.coancestry_curves(W["ploidy"], ids, sl, rl, W["bw"], G, W["grid_start"], K, dl, rec_labels, W["weight_min"],
                                  W["cutoff"], tw, S)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
        if first is None:
            first = m
    return ((pairs.value - p0) if pairs is not None else 0), times, first


def _make_sample_files(tmp, n_inds, tag, seed_shift=0):
    from fastglobetrotter_b200 import synth
    spec = spec_for(n_inds, seed=20261019 + 2 + seed_shift)
    spec.name = f"cfg2_{tag}"
    data = synth.write_dataset(spec, tmp)
    return data.samples_list, data.recom_list, data.rec_labels


def cpu_baseline(sample_inds=12, procs=1, steps=1, warmup=0, tmp=None):
    """the compiled reference companion (hooked build, for its pair counter) on host cores: `procs` processes, each
    reading its own text .gz sample of the workload (`sample_inds` individuals) `warmup + steps` times; only the last
    `steps` calls are timed.  With `tmp` given the files stay there and (files of process 0, means of its first call)
    are returned as well, so that the GPU library can be run on the very same files."""
    from oracle import build_ref
    import multiprocessing as mp
    if not build_ref.build():
        r = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "compiled reference unavailable"}
        return (r, None, None) if tmp else r
    libp = build_ref.HOOKED
    own = None
    if tmp is None:
        own = tempfile.TemporaryDirectory()
        tmp = own.name
    jobs = []
    for p in range(procs):
        sl, rl, rec = _make_sample_files(os.path.join(tmp, f"p{p}"), sample_inds, f"p{p}", seed_shift=p)
        jobs.append((libp, sl, rl, sample_inds, rec, steps, warmup))
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    if procs == 1:
        res = [_ref_worker(jobs[0])]
    else:
        with ctx.Pool(procs) as pool:
            res = pool.map(_ref_worker, jobs)
    wall = time.perf_counter() - t0
    pairs = sum(r[0] for r in res)
    busy = max(sum(r[1]) for r in res)
    out = {"value": pairs / busy, "unit": UNIT, "cores": procs, "kind": "reference", "value_per_core": pairs / busy / procs,
           "sample": f"{procs} process(es) x {steps} timed data_read call(s) (after {warmup} untimed) over {sample_inds} individuals "
                     f"each of the same workload (text .gz inputs, gzip inflate and sscanf parse included, as the reference always "
                     f"does); {pairs} pairs in {busy:.2f} s (wall {wall:.2f} s incl. process start)",
           "pairs": pairs, "seconds": busy, "ms_per_call": busy / max(1, steps) * 1e3}
    if own is not None:
        own.cleanup()
        return out
    return out, (jobs[0][1], jobs[0][2], jobs[0][4]), res[0][2]


def file_parity_and_e2e(lib, files, means_ref, sample_inds, ref_ms_per_call):
    """The R-facing entry point (`data_read`) of the GPU library on THE SAME text .gz files the reference leg just read:
    (1) parity: same srand(1), same files => the curves must equal the compiled reference's bit for bit;
    (2) like-for-like end-to-end time: first call (inflate + tokenise + upload + compute) and cached calls."""
    import ctypes as C
    from fastglobetrotter_b200 import synth
    W = WORK
    sl, rl, rec = files
    K, S, G = W["K"], W["S"], W["G"]
    tw = lib.temp_weights_for_data_read(K, S, synth.random_weights(K, S))
    ids = np.repeat(np.arange(sample_inds, dtype=np.int32), W["n_chrom"])
    dl = spec_for(sample_inds).donor_label_vec()
    args = (W["ploidy"], ids, sl, rl, W["bw"], G, W["grid_start"], K, dl, rec, W["weight_min"], W["cutoff"], tw, S)
    lib.set_option("rand_libc", 1)
    C.CDLL(None).srand(1)
    t0 = time.perf_counter(); m = lib.coancestry_curves(*args); cold = time.perf_counter() - t0
    pairs = lib.last_stats()["pairs"]
    same = bool(means_ref is not None and np.array_equal(m, means_ref))
    warm = []
    for _ in range(3):
        t0 = time.perf_counter(); lib.coancestry_curves(*args); warm.append(time.perf_counter() - t0)
    lib.set_option("rand_libc", 0)
    cached = float(np.median(warm))
    out = {"api": ".C data_read on the reference leg's own text .gz files (process 0's sample)", "individuals": sample_inds,
           "pairs_per_call": pairs, "first_call_ms": cold * 1e3, "cached_call_ms": cached * 1e3,
           "value_first_call": pairs / cold, "value_cached": pairs / cached, "reference_ms_per_call": ref_ms_per_call,
           "speedup_first_call": ref_ms_per_call / (cold * 1e3) if ref_ms_per_call else None,
           "speedup_cached": ref_ms_per_call / (cached * 1e3) if ref_ms_per_call else None}
    # the NULL curves of the first 8 of those individuals: compiled reference vs the GPU library, same files, bit for bit
    try:
        from oracle import build_ref
        from fastglobetrotter_b200.companion import Companion
        ref = Companion(build_ref.HOOKED)
        nn = min(8, sample_inds)
        nargs = (W["ploidy"], sl, rl, W["bw"], G, W["grid_start"], K, dl, rec[:nn], W["weight_min"], W["cutoff"])
        t0 = time.perf_counter(); n_ref = ref.coancestry_curves_null(*nargs); t_ref = time.perf_counter() - t0
        lib.coancestry_curves_null(*nargs)
        t0 = time.perf_counter(); n_gpu = lib.coancestry_curves_null(*nargs); t_gpu = time.perf_counter() - t0
        out["null"] = {"individuals": nn, "pair_evaluations": lib.last_stats()["pairs"], "reference_ms": t_ref * 1e3, "cached_call_ms": t_gpu * 1e3,
                       "parity_checked": bool(np.array_equal(n_ref, n_gpu))}
        same = same and out["null"]["parity_checked"]
    except Exception as e:                                   # the headline must not die on the side check
        out["null"] = {"error": repr(e)}
    return same, out


def other_paths(lib, chrom_dev, common, tail, S, tw, N, dl, K, G, W, steps=3, null_inds=100):
    """the two other entry points of the path on the same device-resident synthetic paintings (outside the timed region of
    the headline): data_read_mode3 (window +1, Y/8, no target-label skip, projection per chromosome) in the headline's
    unit, and data_read_null (exhaustive chunk x chunk across the first `null_inds` individuals) in chunk-pair
    evaluations per second."""
    import torch
    res = {}
    m3 = (3,) + tuple(common[1:])
    dl3 = np.where(dl == 0, 1, dl).astype(np.int32)          # mode 3 never sees the target population label
    tail3 = tail[:5] + (dl3,) + tail[6:]
    lib.data_read_packed(*m3, chrom_dev, *tail3, S, tw)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pairs = 0
    for _ in range(steps):
        _, _, st = lib.data_read_packed(*m3, chrom_dev, *tail3, S, tw)
        pairs += st["pairs"]
    dt = time.perf_counter() - t0
    res["data_read_mode3"] = {"value": pairs / dt, "unit": UNIT, "ms_per_step": dt / steps * 1e3, "pairs_per_step": pairs / steps}
    rows = np.arange(min(null_inds, N), dtype=np.int32)
    lib.data_read_null_packed(W["ploidy"], W["nsamples"], chrom_dev, rows, W["bw"], G, W["grid_start"], K, dl, W["weight_min"], W["cutoff"])
    t0 = time.perf_counter()
    _, st = lib.data_read_null_packed(W["ploidy"], W["nsamples"], chrom_dev, rows, W["bw"], G, W["grid_start"], K, dl, W["weight_min"],
                                      W["cutoff"])
    dt = time.perf_counter() - t0
    peak, peak_src = peaks()
    res["data_read_null"] = {"value": st["pairs"] / dt, "unit": "chunk-pair evaluations/s", "ms": dt * 1e3, "individuals": int(rows.size),
                             "pair_evaluations": st["pairs"], "accepted_fraction": st["accepted"] / max(1, st["pairs"]),
                             "impl": f"null_impl={int(lib.get_option('null_impl'))} (0 = automatic: two phases unless K >= 100 on long chromosomes); "
                                     f"{st['kernel_launches']} kernel launches",
                             "roofline": {"bound": "hbm", "achieved": 16.0 * st["accepted"] / dt / 1e9, "peak": peak, "unit": "GB/s",
                                          "frac": 16.0 * st["accepted"] / dt / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                                          "note": "16 B per accepted update (SURVEY 8d) over the whole call (count, scan, scatter, commit)"}}
    return res


# ------------------------------------------------------------------------------------------------
# BASELINE.json configs[2..4]: biobank-shaped runs (device-generated run-length-encoded paintings), one JSON line each.
# Not part of the default run (the headline stays on configs[1]); `python bench.py --config 3|4|5 [--scale f --chrom-frac f]`.
# ------------------------------------------------------------------------------------------------
BIG = {
    3: dict(name="configs[2]: 22 autosomes, 1,000 target haps (500 diploid inds) x 10 samples, 500k SNPs/chr, K=50, 500 bins, "
                 "one chromosome-bootstrap replicate (R:2058-2059 id matrix), data_read (mode 1)",
            n_chrom=22, n_inds=500, n_snps=500_000, K=50, S=10, mode=1, kind="curves", boot=True),
    4: dict(name="configs[3]: 22 chr, K=100 surrogates, full NULL-individual curve set (data_read_null over the first 100 "
                 "individuals, one sample each), 100k SNPs/chr (~600 chunks per painting)",
            n_chrom=22, n_inds=100, n_snps=100_000, K=100, S=10, mode=1, kind="null", boot=False),
    5: dict(name="configs[4]: one GPU's shard (1/8) of the biobank run: 1,250 of 10,000 diploid inds x 10 samples, 1M SNPs/chr, "
                 "K=150, 500 bins, data_read_mode3 (mode-3 dataflow)",
            n_chrom=22, n_inds=1250, n_snps=1_000_000, K=150, S=10, mode=3, kind="curves", boot=False),
}


def run_big(args):
    import torch
    from fastglobetrotter_b200 import build as fbuild, synth
    from fastglobetrotter_b200.companion import PackedCompanion
    from fastglobetrotter_b200.synth_torch import make_chromosome_runs_device
    cfg = dict(BIG[args.config])
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path")
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    lib = PackedCompanion(fbuild.build())
    lib.set_option("rand_libc", 0)
    for kv in filter(None, os.environ.get("FGT_OPTS", "").split(",")):
        k, v = kv.split("=")
        lib.set_option(k.strip(), float(v))
    n_chrom = max(1, int(round(cfg["n_chrom"] * args.chrom_frac)))
    n_inds = max(2, int(round(cfg["n_inds"] * args.scale)))
    K, S, G, mode = cfg["K"], cfg["S"], 500, cfg["mode"]
    spec = synth.SynthSpec(n_chrom=n_chrom, n_inds=n_inds, ploidy=2, nsamples=10, n_snps=cfg["n_snps"], K=K, haps_per_pop=20,
                           seed=20261019 + args.config, name=f"cfg{args.config}")
    t0 = time.perf_counter()
    chrom, keep, n_runs = [], [], 0
    for h in range(n_chrom):
        pos, rate, run_off, runs = make_chromosome_runs_device(spec, h, dev)
        keep.append(runs)
        n_runs += int(run_off[-1])
        chrom.append({"pos": pos, "rate": rate, "runs": (run_off, int(runs.data_ptr()))})
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t0
    dl = spec.donor_label_vec()
    if mode == 3:
        dl = np.where(dl == 0, 1, dl).astype(np.int32)
    tw = lib.temp_weights_for_data_read(K, S, synth.random_weights(K, S))
    rng = np.random.default_rng(5)
    if cfg["boot"]:
        ids = np.ascontiguousarray(np.sort(rng.integers(0, n_inds, size=(n_chrom, n_inds)), axis=1).T.ravel().astype(np.int32))
    else:
        ids = np.repeat(np.arange(n_inds, dtype=np.int32), n_chrom)
    peak, peak_src = peaks()
    sampler = ClockSampler(0)
    out = {"metric": METRIC, "unit": UNIT, "n_gpus": 1, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
           "data": "synthetic (device-generated run-length-encoded paintings)", "steps": args.steps, "warmup": 1}
    lib.srand(1)
    if cfg["kind"] == "curves":
        call = lambda: lib.data_read_packed(mode, 2, 10, chrom, ids, 0.1, G, 10, K, dl, 0.0, 1.0, S, tw)
        call()                                                 # warm-up (buffers, tables)
        torch.cuda.synchronize()
        sampler.start()
        t0 = time.perf_counter()
        pairs = acc = 0
        ms_accum = launches = chunks = 0
        for _ in range(args.steps):
            m, _, st = call()
            pairs += st["pairs"]; acc += st["accepted"]; ms_accum += st["ms_accum"]; launches += st["accum_launches"]; chunks += st["chunks"]
        dt = time.perf_counter() - t0
        clocks = sampler.stop()
        P = K * (K + 1) // 2
        alg = 16.0 * pairs + 32.0 * chunks + 8.0 * P * G * n_inds * (n_chrom if mode == 3 else 1) * args.steps
        kern_s = ms_accum * 1e-3
        out.update({"value": pairs / dt, "ms_per_step": dt / args.steps * 1e3,
                    "config": {"workload": cfg["name"], "run_at": f"{n_inds} individuals x {n_chrom} chromosomes (scale {args.scale}, chromosome fraction {args.chrom_frac})",
                               "pairs_per_step": pairs / args.steps, "accepted_fraction": acc / max(1, pairs), "runs": n_runs,
                               "paintings_bytes": 8 * n_runs, "generation_s": t_gen, "options": os.environ.get("FGT_OPTS", "")},
                    "roofline": {"bound": "hbm", "kernel": "k_eval_pairs (+ k_route) + k_commit_ordered (all batches of the call)", "achieved": alg / kern_s / 1e9,
                                 "peak": peak, "unit": "GB/s", "frac": alg / kern_s / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                                 "kernel_ms": kern_s * 1e3 / args.steps, "accum_launches_per_step": launches / args.steps},
                    "gpu_launches": None, "clocks": clocks, "means_checksum": float(np.nansum(m))})
    else:
        rows = np.arange(min(100, n_inds), dtype=np.int32)
        call = lambda: lib.data_read_null_packed(2, 10, chrom, rows, 0.1, G, 10, K, dl, 0.0, 1.0)
        call()
        torch.cuda.synchronize()
        sampler.start()
        t0 = time.perf_counter()
        ev = acc = 0
        for _ in range(args.steps):
            t, st = call()
            ev += st["pairs"]; acc += st["accepted"]
        dt = time.perf_counter() - t0
        clocks = sampler.stop()
        out.update({"metric": "null_chunk_pair_evaluations_per_sec", "unit": "chunk-pair evaluations/s", "value": ev / dt,
                    "ms_per_step": dt / args.steps * 1e3,
                    "config": {"workload": cfg["name"], "run_at": f"{rows.size} individuals x {n_chrom} chromosomes (chromosome fraction {args.chrom_frac})",
                               "evaluations_per_step": ev / args.steps, "accepted_fraction": acc / max(1, ev), "runs": n_runs,
                               "options": os.environ.get("FGT_OPTS", "")},
                    "roofline": {"bound": "hbm", "kernel": "data_read_null (two phases: k_null_eval2 count/scatter + k_commit_ordered)", "achieved": 16.0 * acc / dt / 1e9, "peak": peak, "unit": "GB/s",
                                 "frac": 16.0 * acc / dt / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                                 "note": "16 B per accepted update (SURVEY 8d); the tensor lives in shared memory, not HBM"},
                    "clocks": clocks, "checksum": float(np.nansum(t))})
    print(json.dumps(out))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    procs = max(1, min(os.cpu_count() or 1, 32))        # every host core the box has (the reference itself is single-threaded)
    # each step: every process runs one bounded data_read (text .gz sample of the workload); warm-up calls are not timed
    r = cpu_baseline(sample_inds=args.cpu_sample_inds, procs=procs, steps=args.steps, warmup=args.warmup)
    if r["value"] is None:
        print(json.dumps({"impl": "reference", "unavailable": r["sample"]}))
        return
    value = r["value"]
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": r["seconds"] / max(1, args.steps) * 1e3, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic (same generator, text .gz files)",
           "config": {"workload": WORKLOAD_NAME, "sample": r["sample"]},
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "reference", "sample": r["sample"],
                            "value_per_core": r["value_per_core"]},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-dense", action="store_true", help="skip the dense-label-matrix comparison numbers")
    ap.add_argument("--variants", action="store_true", help="also time the kernel variants (bulk rand stream, lane-queue commit)")
    ap.add_argument("--cpu-sample-inds", type=int, default=16)
    ap.add_argument("--config", type=int, default=0, choices=[0, 3, 4, 5], help="run one of BASELINE.json configs[2..4] (numbered 3..5) instead of the headline")
    ap.add_argument("--scale", type=float, default=1.0, help="--config: fraction of the configuration's individuals to run")
    ap.add_argument("--chrom-frac", type=float, default=1.0, help="--config: fraction of the 22 chromosomes to run")
    args = ap.parse_args()
    if args.config and args.impl != "reference":
        if args.steps == 10:
            args.steps = 1
        run_big(args)
        return
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
