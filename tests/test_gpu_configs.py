"""GPU parity on BASELINE.json's own configurations.

configs[0]  the reference's tutorial IN FULL (100 individuals x 10 samples, chr21 + chr22, K = 26, 291 bins) through the
            R-facing .C symbols on the text files, against what the compiled reference produced here
            (tests/golden/tutorial_full.npz, made by tests/golden/make_golden.py): raw tensors, curves, NULL tensor.
configs[1]  at its real size per individual (100k SNPs, K = 30, 500 bins) on 16 individuals against the COMPILED REFERENCE
            run on the spot (oracle/_ref travels with gpurun), raw tensors and curves, modes 1 and 3.
configs[2..4] scaled slices against the oracle: K = 50 / 100 / 150, 500 bins, 22 chromosomes, a 1,000,000-SNP chromosome,
            the NULL tensor at K = 100 -- the shapes on which the large-K (global tensor) path and the long schedules run.
Everything is bit-exact (np.array_equal on float64)."""
import ctypes as C
import os

import numpy as np
import pytest

import golden_util as gu
from helpers import GOLDEN, REF_HOOKED, bootstrap_ids, have_ref, id_file_inputs, ind_id_matrix, parse_paramfile, srand
from fastglobetrotter_b200 import synth

pytestmark = pytest.mark.gpu


def test_full_tutorial_through_the_C_entry_points(lib):
    gold = gu.load("tutorial_full.npz")
    tdir = os.path.join(GOLDEN, "tutorial")
    prm = parse_paramfile(os.path.join(tdir, "paramfile.txt"))
    donors = prm["copyvector.popnames"].split()
    dl, rec = id_file_inputs(os.path.join(tdir, "individual.txt"), donors, prm["target.popname"])
    assert np.array_equal(dl, gold["donor_label_vec"]) and len(rec) == int(gold["N"]) == 100
    K, S, bw, G, gs, stride = int(gold["K"]), int(gold["S"]), float(gold["bw"]), int(gold["G"]), int(gold["gs"]), int(gold["stride"])
    tw = gold["tw"]
    cwd = os.getcwd()
    os.chdir(GOLDEN)
    lib.set_option("rand_libc", 1)
    lib.set_option("keep_raw", 1)
    try:
        for tag, mode in (("", 1), ("", 3), ("_boot", 1)):
            ids = gold[f"ids{tag}"]
            srand(int(gold["seed"]))
            f = lib.coancestry_curves if mode == 1 else lib.coancestry_curves_mode3
            m = f(2, ids, "tutorial/samplefile.txt", "tutorial/recomfile.txt", bw, G, gs, K, dl, rec, 0.0, 1.0, tw, S)
            st = lib.last_stats()
            assert st["pairs"] == int(gold[f"pairs{mode}{tag}"]), (mode, tag)           # 131,353,741 for data_read
            gu.assert_matches(gold, f"raw{mode}{tag}", lib.get_raw_counts(K, G), stride)
            assert np.array_equal(m, gold[f"means{mode}{tag}"]), (mode, tag)
        lib.set_option("keep_raw", 0)
        nul = lib.coancestry_curves_null(2, "tutorial/samplefile.txt", "tutorial/recomfile.txt", bw, G, gs, K, dl, rec, 0.0, 1.0)
        gu.assert_matches(gold, "null", nul, stride)
    finally:
        os.chdir(cwd)
        lib.set_option("rand_libc", 0)
        lib.set_option("keep_raw", 0)


@pytest.mark.skipif(not have_ref(), reason="compiled reference (oracle/_ref) not present")
@pytest.mark.parametrize("mode", [1, 3])
def test_config1_size_against_the_compiled_reference(lib, ref_libs, tmp_path, mode):
    """BASELINE.json configs[1] at full size per individual: 100,000 SNPs, 10 samples, K = 30, 500 bins of 0.1 cM, S = 10,
    16 individuals; the same text .gz files and the same srand() for the compiled reference and the GPU library."""
    _, hooked = ref_libs
    n_inds = 16
    spec = synth.SynthSpec(n_chrom=1, n_inds=n_inds, ploidy=2, nsamples=10, n_snps=100_000, K=30, haps_per_pop=20, seed=20261019 + 2,
                           name="cfg2_parity")
    if mode == 3:
        spec.self_copy_frac = 0.0
    data = synth.write_dataset(spec, str(tmp_path))
    K, S, G, bw, gs = 30, 10, 500, 0.1, 10
    tw = hooked.temp_weights_for_data_read(K, S, synth.random_weights(K, S))
    ids = ind_id_matrix(n_inds, 1)
    slots = n_inds
    raw_r = np.zeros((slots, K * K, G))
    sym = "REF_DUMP_counts" if mode == 1 else "REF_DUMP_counts3"
    for nm in ("REF_DUMP_counts", "REF_DUMP_counts3"):
        C.c_void_p.in_dll(hooked.lib, nm).value = None
    C.c_void_p.in_dll(hooked.lib, sym).value = raw_r.ctypes.data
    C.c_long.in_dll(hooked.lib, "REF_DUMP_calls3").value = 0
    C.c_long.in_dll(hooked.lib, "REF_DUMP_pairs").value = 0
    args = (2, ids, data.samples_list, data.recom_list, bw, G, gs, K, data.donor_label_vec, data.rec_labels, 0.0, 1.0, tw, S)
    srand(31)
    m_r = (hooked.coancestry_curves if mode == 1 else hooked.coancestry_curves_mode3)(*args)
    pairs_r = C.c_long.in_dll(hooked.lib, "REF_DUMP_pairs").value
    C.c_void_p.in_dll(hooked.lib, sym).value = None
    lib.set_option("rand_libc", 1)
    lib.set_option("keep_raw", 1)
    try:
        srand(31)
        m_g = (lib.coancestry_curves if mode == 1 else lib.coancestry_curves_mode3)(*args)
        st = lib.last_stats()
        raw_g = lib.get_raw_counts(K, G)
    finally:
        lib.set_option("rand_libc", 0)
        lib.set_option("keep_raw", 0)
    assert st["pairs"] == pairs_r and pairs_r > 20_000_000
    assert np.array_equal(raw_g, raw_r), int((raw_g != raw_r).sum())
    assert np.array_equal(m_g, m_r)


def _chroms(spec):
    out = []
    for h in range(spec.n_chrom):
        pos, rate, labels = synth.make_chromosome(spec, h)
        out.append({"pos": pos, "rate": rate, "labels": labels})
    return out


def _both(lib, oracle, mode, spec, chrom, ids, G, S, seed, K=None):
    K = K or spec.K
    dl = spec.donor_label_vec()
    tw = np.random.default_rng(seed).random(K * K * S * S)
    oracle.srand(seed)
    m_o, raw_o, pairs_o = oracle.data_read_packed(mode, spec.ploidy, spec.nsamples, chrom, ids, 0.1, G, 10, K, dl, 0.0, 1.0, S, tw, want_raw=True)
    lib.set_option("rand_libc", 0)
    lib.srand(seed)
    m_g, raw_g, st = lib.data_read_packed(mode, spec.ploidy, spec.nsamples, chrom, ids, 0.1, G, 10, K, dl, 0.0, 1.0, S, tw, want_raw=True)
    assert st["pairs"] == pairs_o
    assert np.array_equal(raw_g, raw_o), int((raw_g != raw_o).sum())
    assert np.array_equal(m_g, m_o)
    return st


def test_config3_slice_22_chromosomes_K50_with_bootstrap_ids(lib, oracle):
    """configs[2] scaled: 22 chromosomes of different lengths (35..150 cM, so the 53-cM window both fits and does not),
    K = 50 donor groups, 500 bins, a bootstrap id matrix (R:2058-2059), both modes."""
    total = 0
    for mode in (1, 3):
        spec = synth.SynthSpec(n_chrom=22, n_inds=3, ploidy=2, nsamples=6, n_snps=2400, K=50, haps_per_pop=6, rate=6e-7, seed=303 + mode,
                               name="cfg3")
        chrom = _chroms(spec)
        for h, c in enumerate(chrom):                      # chromosome h keeps the first share of the SNPs: lengths 35 .. 150 cM
            keep = int(2400 * (1.0 - 0.035 * h))
            c["pos"], c["rate"], c["labels"] = c["pos"][:keep].copy(), c["rate"][:keep].copy(), np.ascontiguousarray(c["labels"][:, :keep])
            c["rate"][-1] = 0.0
        st = _both(lib, oracle, mode, spec, chrom, bootstrap_ids(3, 22, 5), 500, 5, 100 + mode)
        total += st["pairs"]
    assert total > 3_000_000


def test_config4_slice_K100_curves_and_full_null(lib, oracle, each_null_impl):
    """configs[3] scaled: K = 100 surrogates/donor groups (tile too large for shared memory: global-tensor commit path),
    500 bins, three chromosomes; data_read and the complete NULL-individual tensor (100 x 100 x 500) over 5 individuals."""
    spec = synth.SynthSpec(n_chrom=3, n_inds=5, ploidy=2, nsamples=5, n_snps=9000, K=100, haps_per_pop=4, rate=1.4e-7, seed=404, name="cfg4")
    chrom = _chroms(spec)
    st = _both(lib, oracle, 1, spec, chrom, ind_id_matrix(5, 3), 500, 6, 44)
    assert st["pairs"] > 1_000_000
    dl = spec.donor_label_vec()
    n_o, ev = oracle.data_read_null_packed(2, 5, chrom, np.arange(5), 0.1, 500, 10, 100, dl, 0.0, 1.0)
    n_g, st = lib.data_read_null_packed(2, 5, chrom, np.arange(5), 0.1, 500, 10, 100, dl, 0.0, 1.0)
    assert st["pairs"] == ev and ev > 5_000_000
    assert np.array_equal(n_g, n_o)


def test_config5_slice_K150_one_million_snps_mode3(lib, oracle):
    """configs[4] scaled: one chromosome of 1,000,000 SNPs (1,500 cM: 1,501 coarse bins), K = 150, 500 bins, mode 3
    (the only dataflow that scales), two individuals x 10 samples."""
    spec = synth.SynthSpec(n_chrom=1, n_inds=2, ploidy=2, nsamples=10, n_snps=1_000_000, K=150, haps_per_pop=3, seed=505, name="cfg5")
    chrom = _chroms(spec)
    st = _both(lib, oracle, 3, spec, chrom, ind_id_matrix(2, 1), 500, 4, 55)
    assert st["pairs"] > 25_000_000
