"""CPU tests of the oracle (oracle/companion_oracle.c): against libc, against the golden vectors
generated from the compiled reference, and -- where /root/reference is present -- against the
compiled reference itself on fresh synthetic data and on the tutorial."""
import ctypes as C
import os

import numpy as np
import pytest

import golden_util as gu
from helpers import (REF_HOOKED, TUTORIAL, bootstrap_ids, have_ref, id_file_inputs, ind_id_matrix, libc_rand, parse_paramfile,
                     srand)
from fastglobetrotter_b200 import synth


def test_rand_emulator_matches_libc(oracle):
    for seed in (1, 12345, 0, 2**31 + 7):
        srand(seed)
        oracle.srand(seed)
        want = [libc_rand() for _ in range(100000)]
        got = [oracle.rand() for _ in range(100000)]
        assert want == got, seed
    oracle.srand(1)
    assert [oracle.rand() for _ in range(4)] == [1804289383, 846930886, 1681692777, 1714636915]


def test_oracle_vs_golden_tutorial(oracle):
    g = gu.load("tutorial_subset.npz")
    chrom = gu.tutorial_chrom(g)
    K, S, G, bw, gs, N = (int(g[k]) for k in ("K", "S", "G", "bw", "gs", "N")) if False else \
        (int(g["K"]), int(g["S"]), int(g["G"]), float(g["bw"]), int(g["gs"]), int(g["N"]))
    for mode in (1, 3):
        oracle.srand(int(g["seed"]))
        m, raw, pairs = oracle.data_read_packed(mode, 2, 10, chrom, g["ids"], bw, G, gs, K, g["donor_label_vec"], 0.0, 1.0, S,
                                                g["tw"], want_raw=True)
        assert pairs == int(g[f"pairs{mode}"])
        assert np.array_equal(m, g[f"means{mode}"])
        gu.assert_matches(g, f"raw{mode}", raw)
    null, _ = oracle.data_read_null_packed(2, 10, chrom, np.arange(N), bw, G, gs, K, g["donor_label_vec"], 0.0, 1.0)
    gu.assert_matches(g, "null", null)


def test_oracle_vs_golden_synth(oracle):
    g = gu.load("synth_a.npz")
    spec, chrom, dl = gu.synth_a_packed()
    K, S, G, bw, gs = int(g["K"]), int(g["S"]), int(g["G"]), float(g["bw"]), int(g["gs"])
    oracle.srand(int(g["seed"]))
    m, raw, pairs = oracle.data_read_packed(1, spec.ploidy, spec.nsamples, chrom, g["ids"], bw, G, gs, K, dl, 0.0, 1.0, S,
                                            g["tw"], want_raw=True)
    assert pairs == int(g["pairs1"])
    assert np.array_equal(m, g["means1"])
    gu.assert_matches(g, "raw1", raw)
    null, _ = oracle.data_read_null_packed(spec.ploidy, spec.nsamples, chrom, np.arange(spec.n_inds), bw, G, gs, K, dl, 0.0, 1.0)
    gu.assert_matches(g, "null", null)


def _packed(data):
    sp = data.spec
    per = sp.ploidy * sp.nsamples
    out = []
    for c in data.chrom:
        rows = np.concatenate([np.arange(s * per, (s + 1) * per) for s in data.rec_file_index])
        out.append({"pos": c["pos"], "rate": c["rate"], "labels": np.ascontiguousarray(c["labels"][rows])})
    return out


@pytest.mark.skipif(not have_ref(), reason="compiled reference not available")
@pytest.mark.parametrize("mode", [1, 3])
@pytest.mark.parametrize("case", ["plain", "bootstrap_extra", "weight_min"])
def test_oracle_vs_compiled_reference_synthetic(oracle, ref_libs, tmp_path, mode, case):
    _, ref = ref_libs
    kw = dict(n_chrom=2, n_inds=5, n_snps=2500, K=6, nsamples=3, rate=5e-7, seed=99 + mode)
    if case == "bootstrap_extra":
        kw.update(n_extra_inds=3, self_copy_frac=0.0 if mode == 3 else 0.05)
    spec = synth.SynthSpec(name=f"t_{case}_{mode}", **kw)
    data = synth.write_dataset(spec, str(tmp_path))
    K, S, G, bw, gs = spec.K, 3, 90, 0.1, 10
    wmin = 2.0 if case == "weight_min" else 0.0
    W = synth.random_weights(K, S)
    tw = ref.temp_weights_for_data_read(K, S, W)
    assert np.array_equal(tw, oracle.mutiply_temp_weight(K, S, W.ravel(order="F")).reshape(S * S, K * K).ravel(order="F"))
    ids = bootstrap_ids(spec.n_inds, spec.n_chrom, 5) if case == "bootstrap_extra" else ind_id_matrix(spec.n_inds, spec.n_chrom)
    n_slots = spec.n_inds if mode == 1 else spec.n_inds * spec.n_chrom
    raw_ref = np.zeros((n_slots, K * K, G))
    sym = "REF_DUMP_counts" if mode == 1 else "REF_DUMP_counts3"
    C.c_void_p.in_dll(ref.lib, "REF_DUMP_counts").value = None
    C.c_void_p.in_dll(ref.lib, "REF_DUMP_counts3").value = None
    C.c_void_p.in_dll(ref.lib, sym).value = raw_ref.ctypes.data
    C.c_long.in_dll(ref.lib, "REF_DUMP_calls3").value = 0
    C.c_long.in_dll(ref.lib, "REF_DUMP_pairs").value = 0
    srand(4242)
    f = ref.coancestry_curves if mode == 1 else ref.coancestry_curves_mode3
    m_ref = f(spec.ploidy, ids, data.samples_list, data.recom_list, bw, G, gs, K, data.donor_label_vec, data.rec_labels, wmin,
              0.8, tw, S)
    pairs_ref = C.c_long.in_dll(ref.lib, "REF_DUMP_pairs").value
    C.c_void_p.in_dll(ref.lib, sym).value = None
    oracle.srand(4242)
    m_o, raw_o, pairs_o = oracle.data_read_packed(mode, spec.ploidy, spec.nsamples, _packed(data), ids, bw, G, gs, K,
                                                  data.donor_label_vec, wmin, 0.8, S, tw, want_raw=True)
    assert pairs_o == pairs_ref
    assert np.array_equal(raw_o, raw_ref)
    assert np.array_equal(m_o, m_ref)
    # the emulated stream is where libc's is
    assert [oracle.rand() for _ in range(3)] == [libc_rand() for _ in range(3)]
    if mode == 1:
        n_ref = ref.coancestry_curves_null(spec.ploidy, data.samples_list, data.recom_list, bw, G, gs, K, data.donor_label_vec,
                                           data.rec_labels, wmin, 0.8)
        n_o, _ = oracle.data_read_null_packed(spec.ploidy, spec.nsamples, _packed(data), np.arange(spec.n_inds), bw, G, gs, K,
                                              data.donor_label_vec, wmin, 0.8)
        assert np.array_equal(n_o, n_ref)


@pytest.mark.skipif(not (have_ref() and os.path.isdir(TUTORIAL)), reason="needs /root/reference")
def test_golden_files_are_what_the_reference_produces_today(ref_libs):
    """re-run the compiled reference on the tutorial subset and compare with the committed fixture"""
    _, ref = ref_libs
    g = gu.load("tutorial_subset.npz")
    prm = parse_paramfile(os.path.join(TUTORIAL, "paramfile.txt"))
    dl, rec = id_file_inputs(os.path.join(TUTORIAL, "individual.txt"), prm["copyvector.popnames"].split(), prm["target.popname"])
    assert np.array_equal(dl, g["donor_label_vec"])
    cwd = os.getcwd()
    os.chdir(os.path.dirname(TUTORIAL))
    try:
        srand(int(g["seed"]))
        m = ref.coancestry_curves(2, g["ids"], "tutorial/samplefile.txt", "tutorial/recomfile.txt", float(g["bw"]), int(g["G"]),
                                  int(g["gs"]), int(g["K"]), dl, rec[:int(g["N"])], 0.0, 1.0, g["tw"], int(g["S"]))
    finally:
        os.chdir(cwd)
    assert np.array_equal(m, g["means1"])
