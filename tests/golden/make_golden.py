#!/usr/bin/env python3
"""Generates tests/golden/*.npz from the COMPILED REFERENCE (oracle/_ref, built by
oracle/build_ref.py from /root/reference) run in this container.  Commit the outputs; the
GPU box has no /root/reference.

  tutorial_full.npz    ALL 100 target individuals of the reference's tutorial (chr21+chr22, BASELINE.json configs[0]):
                       what the compiled reference computes through data_read, data_read_mode3 (plain and with a
                       bootstrap id matrix) and data_read_null on the text files committed under tests/golden/tutorial/
                       (copies of the reference's tutorial inputs: test vectors); outputs as sha256 + strided samples.
  tutorial_subset.npz  the first 2 target individuals of the reference's tutorial (chr21+chr22),
                       re-packed (uint16 donor-haplotype ids + recombination maps) together with
                       what the reference computes from them through its .C entry points.
  synth_a.npz          outputs of the reference on a synthetic data set that the tests regenerate
                       from its seed (fastglobetrotter_b200.synth), so only outputs are stored.

Stored per case: `means` for data_read and data_read_mode3 (libc srand seed given), the NULL tensor,
the number of sampled pairs, and the raw per-individual tensors (total_counts_mat_all, through the
hooked build) -- everything bit-exact (float64, lossless npz).
"""
import ctypes as C
import gzip
import hashlib
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import REF_HOOKED, TUTORIAL, id_file_inputs, ind_id_matrix, bootstrap_ids, parse_paramfile, srand  # noqa: E402
from fastglobetrotter_b200 import synth  # noqa: E402
from fastglobetrotter_b200.companion import Companion  # noqa: E402
from oracle import build_ref  # noqa: E402

SEED = 20261019
STRIDE = 53


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.float64).tobytes()).hexdigest()


def compact(out: dict, key: str, a: np.ndarray):
    """big fp64 tensors are stored as sha256 of their bytes + every STRIDE-th value (bit-exact)"""
    out[key + "_sha256"] = digest(a)
    out[key + "_sample"] = np.ascontiguousarray(a.ravel()[::STRIDE])
    out[key + "_shape"] = np.array(a.shape)
    out[key + "_sum"] = float(a.sum())



def run_ref(ref, mode, ploidy, ids, sl, rl, bw, G, gs, K, dl, rec, S, tw, n_calls_dim):
    raw = np.zeros((n_calls_dim, K * K, G))
    sym = "REF_DUMP_counts" if mode == 1 else "REF_DUMP_counts3"
    C.c_void_p.in_dll(ref.lib, "REF_DUMP_counts").value = None
    C.c_void_p.in_dll(ref.lib, "REF_DUMP_counts3").value = None
    C.c_void_p.in_dll(ref.lib, sym).value = raw.ctypes.data
    C.c_long.in_dll(ref.lib, "REF_DUMP_calls3").value = 0
    C.c_long.in_dll(ref.lib, "REF_DUMP_pairs").value = 0
    srand(SEED)
    f = ref.coancestry_curves if mode == 1 else ref.coancestry_curves_mode3
    means = f(ploidy, ids, sl, rl, bw, G, gs, K, dl, rec, 0.0, 1.0, tw, S)
    pairs = C.c_long.in_dll(ref.lib, "REF_DUMP_pairs").value
    C.c_void_p.in_dll(ref.lib, sym).value = None
    return means, raw, pairs


def parse_samples(path, rec, nsamples, ploidy):
    rows = []
    with gzip.open(path, "rt") as f:
        f.readline()
        cur = False
        for ln in f:
            if ln.startswith("HAP"):
                cur = ln.split()[2] in rec
                if not cur and len(rows) == len(rec) * ploidy * nsamples:
                    break
                continue
            if cur:
                rows.append(np.array(ln.split()[1:], dtype=np.int32))
    return rows


def tutorial_case(ref):
    prm = parse_paramfile(os.path.join(TUTORIAL, "paramfile.txt"))
    donors = prm["copyvector.popnames"].split()
    dl, rec_all = id_file_inputs(os.path.join(TUTORIAL, "individual.txt"), donors, prm["target.popname"])
    N = 2
    rec = rec_all[:N]
    K, S, bw, G, gs = len(donors), 4, 0.1, 291, 10
    cwd = os.getcwd()
    os.chdir(os.path.dirname(TUTORIAL))
    try:
        chrom = {}
        sfiles = open("tutorial/samplefile.txt").read().split()
        rfiles = open("tutorial/recomfile.txt").read().split()
        for h, (sp, rp) in enumerate(zip(sfiles, rfiles)):
            pos, rate = synth.read_recomb_file(rp)
            rows = parse_samples(sp, rec, 10, 2)
            lab = np.stack([r[:pos.size] for r in rows]).astype(np.uint16)
            chrom[f"pos{h}"], chrom[f"rate{h}"], chrom[f"labels{h}"] = pos, rate, lab
        W = synth.random_weights(K, S, seed=11)
        tw = ref.temp_weights_for_data_read(K, S, W)
        ids = ind_id_matrix(N, 2)
        out = dict(chrom)
        for mode in (1, 3):
            m, raw, pairs = run_ref(ref, mode, 2, ids, "tutorial/samplefile.txt", "tutorial/recomfile.txt", bw, G, gs, K,
                                    dl, rec, S, tw, N if mode == 1 else 2 * N)
            out[f"means{mode}"], out[f"pairs{mode}"] = m, pairs
            compact(out, f"raw{mode}", raw)
        compact(out, "null", ref.coancestry_curves_null(2, "tutorial/samplefile.txt", "tutorial/recomfile.txt", bw, G, gs, K,
                                                        dl, rec, 0.0, 1.0))
    finally:
        os.chdir(cwd)
    out.update(dict(donor_label_vec=dl, K=K, S=S, bw=bw, G=G, gs=gs, N=N, seed=SEED, tw=tw, ids=ids, n_chrom=2))
    np.savez_compressed(os.path.join(HERE, "tutorial_subset.npz"), **out)
    print("tutorial_subset.npz", {k: v for k, v in out.items() if k.startswith("pairs") or k.endswith("_sum")})


FULL_STRIDE = 997


def compact_full(out: dict, key: str, a: np.ndarray):
    out[key + "_sha256"] = digest(a)
    out[key + "_sample"] = np.ascontiguousarray(a.ravel()[::FULL_STRIDE])
    out[key + "_shape"] = np.array(a.shape)
    out[key + "_sum"] = float(a.sum())


def tutorial_full_case(ref):
    """BASELINE.json configs[0] in full: 100 individuals x 10 samples, chr21 + chr22, K = 26, 291 bins, S = 7."""
    tdir = os.path.join(HERE, "tutorial")
    prm = parse_paramfile(os.path.join(tdir, "paramfile.txt"))
    donors = prm["copyvector.popnames"].split()
    dl, rec = id_file_inputs(os.path.join(tdir, "individual.txt"), donors, prm["target.popname"])
    N = len(rec)
    K, S, bw, G, gs = len(donors), 7, 0.1, 291, 10
    W = synth.random_weights(K, S, seed=13)
    tw = ref.temp_weights_for_data_read(K, S, W)
    cwd = os.getcwd()
    os.chdir(HERE)                     # the list files name their entries relative to the parent of tutorial/
    out = {}
    try:
        for tag, ids in (("", ind_id_matrix(N, 2)), ("_boot", bootstrap_ids(N, 2, 2026))):
            for mode in (1, 3):
                if tag and mode == 3:
                    continue
                m, raw, pairs = run_ref(ref, mode, 2, ids, "tutorial/samplefile.txt", "tutorial/recomfile.txt", bw, G, gs, K,
                                        dl, rec, S, tw, N if mode == 1 else 2 * N)
                out[f"means{mode}{tag}"], out[f"pairs{mode}{tag}"] = m, pairs
                compact_full(out, f"raw{mode}{tag}", raw)
                print("tutorial_full", tag, mode, pairs, flush=True)
            out[f"ids{tag}"] = ids
        compact_full(out, "null", ref.coancestry_curves_null(2, "tutorial/samplefile.txt", "tutorial/recomfile.txt", bw, G, gs, K,
                                                             dl, rec, 0.0, 1.0))
    finally:
        os.chdir(cwd)
    out.update(dict(donor_label_vec=dl, K=K, S=S, bw=bw, G=G, gs=gs, N=N, seed=SEED, tw=tw, n_chrom=2, stride=FULL_STRIDE))
    np.savez_compressed(os.path.join(HERE, "tutorial_full.npz"), **out)
    print("tutorial_full.npz", {k: v for k, v in out.items() if k.startswith("pairs") or k.endswith("_sum")})


def synth_spec():
    return synth.SynthSpec(n_chrom=2, n_inds=5, n_snps=3500, K=8, nsamples=4, self_copy_frac=0.03, n_extra_inds=2, rate=3.5e-7,
                           seed=SEED + 5, name="golden_a")


def synth_case(ref):
    spec = synth_spec()
    K, S, bw, G, gs = spec.K, 3, 0.1, 130, 10
    with tempfile.TemporaryDirectory() as d:
        data = synth.write_dataset(spec, d)
        W = synth.random_weights(K, S, seed=3)
        tw = ref.temp_weights_for_data_read(K, S, W)
        ids = bootstrap_ids(spec.n_inds, spec.n_chrom, 21)
        out = {}
        m, raw, pairs = run_ref(ref, 1, 2, ids, data.samples_list, data.recom_list, bw, G, gs, K, data.donor_label_vec,
                                data.rec_labels, S, tw, spec.n_inds)
        out["means1"], out["pairs1"] = m, pairs
        compact(out, "raw1", raw)
        compact(out, "null", ref.coancestry_curves_null(2, data.samples_list, data.recom_list, bw, G, gs, K,
                                                        data.donor_label_vec, data.rec_labels, 0.0, 1.0))
        out.update(dict(K=K, S=S, bw=bw, G=G, gs=gs, seed=SEED, tw=tw, ids=ids))
    np.savez_compressed(os.path.join(HERE, "synth_a.npz"), **out)
    print("synth_a.npz pairs", pairs)


if __name__ == "__main__":
    assert build_ref.build(), "needs /root/reference"
    ref = Companion(REF_HOOKED)
    which = sys.argv[1:] or ["subset", "synth", "full"]
    if "subset" in which:
        tutorial_case(ref)
    if "synth" in which:
        synth_case(ref)
    if "full" in which:
        tutorial_full_case(ref)
