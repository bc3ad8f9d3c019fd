"""Loading / checking of the golden fixtures produced by tests/golden/make_golden.py."""
import hashlib
import os

import numpy as np

from helpers import GOLDEN

STRIDE = 53


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.float64).tobytes()).hexdigest()


def assert_matches(gold, key, arr, stride=STRIDE):
    arr = np.asarray(arr, dtype=np.float64)
    assert tuple(gold[key + "_shape"]) == arr.shape, (key, arr.shape)
    assert np.array_equal(arr.ravel()[::stride], gold[key + "_sample"]), f"{key}: sampled cells differ"
    assert digest(arr) == str(gold[key + "_sha256"]), f"{key}: sha256 differs"


def tutorial_chrom(gold):
    return [{"pos": gold[f"pos{h}"], "rate": gold[f"rate{h}"], "labels": gold[f"labels{h}"]} for h in range(int(gold["n_chrom"]))]


def synth_a_spec():
    from fastglobetrotter_b200 import synth
    return synth.SynthSpec(n_chrom=2, n_inds=5, n_snps=3500, K=8, nsamples=4, self_copy_frac=0.03, n_extra_inds=2, rate=3.5e-7,
                           seed=20261019 + 5, name="golden_a")


def synth_a_packed():
    """regenerate the packed inputs of synth_a.npz (recipients only, text round trip of the map)"""
    import tempfile
    from fastglobetrotter_b200 import synth
    spec = synth_a_spec()
    with tempfile.TemporaryDirectory() as d:
        data = synth.write_dataset(spec, d)
    per = spec.ploidy * spec.nsamples
    chrom = []
    for c in data.chrom:
        rows = np.concatenate([np.arange(s * per, (s + 1) * per) for s in data.rec_file_index])
        chrom.append({"pos": c["pos"], "rate": c["rate"], "labels": np.ascontiguousarray(c["labels"][rows])})
    return spec, chrom, data.donor_label_vec
