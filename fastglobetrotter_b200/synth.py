"""Synthetic ChromoPainter paintings of the shapes named in BASELINE.json.

There is no network and the only real data set is the reference's tutorial, so every
benchmark/parity configuration other than the tutorial is generated here:

* uniform recombination map (SNP spacing ``bp_step`` bp, ``rate`` Morgans/bp, last row 0),
* a donor panel of ``K`` populations x ``haps_per_pop`` haplotypes,
* every painting (one sampled ChromoPainter path = one donor-haplotype id per SNP) is a
  2-source ancestry-tract process ``gens`` generations old (so that the coancestry
  curves decay instead of being flat) emitting donor-haplotype runs with exponential
  length (mean ``mean_chunk_cm``); the donor population is drawn from the tract's
  source-specific K-vector and the haplotype uniformly inside the population.

Outputs are both *packed* arrays (what the library's packed entry points and the oracle
eat) and the *text* files the reference companion reads (``*.samples.out.gz``, recomb maps,
list files; format per reference C:232-275, C:353-381, C:436-445).
"""
from __future__ import annotations

import gzip
import os
from dataclasses import dataclass, field

import numpy as np


@dataclass
class SynthSpec:
    n_chrom: int = 1
    n_inds: int = 4            # recipient (target) individuals
    ploidy: int = 2
    nsamples: int = 10
    n_snps: int = 2000
    K: int = 6                 # donor populations
    haps_per_pop: int = 20
    bp_step: float = 1000.0
    rate: float = 1.5e-8       # Morgans per bp -> 100k SNPs = 150 cM
    mean_chunk_cm: float = 0.25
    gens: float = 30.0
    self_copy_frac: float = 0.0   # fraction of chunks copied from the target population (label 0)
    n_extra_inds: int = 0         # non-recipient individuals interleaved in the samples file
    seed: int = 20261019
    name: str = "synth"
    rec_labels: list = field(default_factory=list)

    def __post_init__(self):
        if not self.rec_labels:
            self.rec_labels = [f"Target{i + 1}" for i in range(self.n_inds)]

    @property
    def n_donor_haps(self) -> int:
        return self.K * self.haps_per_pop

    @property
    def cm_per_snp(self) -> float:
        return self.rate * self.bp_step * 100.0

    def donor_label_vec(self) -> np.ndarray:
        """What R builds at fastGLOBETROTTER.R:808-820: one population label per haplotype
        of the id file (1..K donors, 0 = target population); donors first, then targets."""
        lab = np.repeat(np.arange(1, self.K + 1, dtype=np.int32), self.haps_per_pop)
        tgt = np.zeros((self.n_inds + self.n_extra_inds) * self.ploidy, dtype=np.int32)
        return np.concatenate([lab, tgt])


def recomb_map(spec: SynthSpec, chrom: int):
    pos = 1.0e6 + spec.bp_step * np.arange(spec.n_snps, dtype=np.float64)
    rng = np.random.default_rng(spec.seed + 7919 * (chrom + 1))
    # mild rate heterogeneity (x0.5 .. x1.5) so that cM positions are not a lattice
    rate = spec.rate * (0.5 + rng.random(spec.n_snps))
    rate[-1] = 0.0
    return pos, rate


def _paint_one(rng, spec: SynthSpec, cm_at_snp: np.ndarray, srcvec: np.ndarray) -> np.ndarray:
    total = cm_at_snp[-1]
    n_snps = spec.n_snps
    # ancestry tracts: switch rate gens/100 per cM
    nt = int(total * spec.gens / 100.0 * 1.6) + 16
    tb = np.cumsum(rng.exponential(100.0 / spec.gens, nt))
    t_src0 = rng.integers(0, 2)
    # chunk starts
    nc = int(total / spec.mean_chunk_cm * 1.4) + 64
    cb = np.concatenate([[0.0], np.cumsum(rng.exponential(spec.mean_chunk_cm, nc))])
    cb = cb[cb < total]
    src = (t_src0 + np.searchsorted(tb, cb, side="right")) & 1
    u = rng.random(cb.size)
    cdf = np.cumsum(srcvec, axis=1)
    pop = np.where(src == 0, np.searchsorted(cdf[0], u), np.searchsorted(cdf[1], u)).clip(0, spec.K - 1)
    hap = pop * spec.haps_per_pop + rng.integers(0, spec.haps_per_pop, cb.size) + 1
    if spec.self_copy_frac > 0:
        selfc = rng.random(cb.size) < spec.self_copy_frac
        hap = np.where(selfc, spec.n_donor_haps + 1 + rng.integers(0, spec.n_inds * spec.ploidy, cb.size), hap)
    start = np.searchsorted(cm_at_snp, cb, side="left")
    start[0] = 0
    lens = np.diff(np.concatenate([start, [n_snps]]))
    return np.repeat(hap, lens)


def cm_positions(pos: np.ndarray, rate: np.ndarray) -> np.ndarray:
    """Approximate cM coordinate per SNP (generator-internal; the bit-exact sequential
    prefix lives in the library and the oracle, reference C:494)."""
    inc = rate[:-1] * np.diff(pos) * 100.0
    return np.concatenate([[0.0], np.cumsum(inc)])


def make_chromosome(spec: SynthSpec, chrom: int):
    """Returns (pos, rate, labels) with labels[(ind*ploidy + p)*nsamples + s, snp] = donor hap id (1-based).
    Individuals here are ALL individuals of the samples file (recipients first unless extras are
    interleaved, see :func:`file_order`)."""
    pos, rate = recomb_map(spec, chrom)
    cm = cm_positions(pos, rate)
    rng = np.random.default_rng(spec.seed + 104729 * (chrom + 1))
    prng = np.random.default_rng(spec.seed + 1)
    srcvec = prng.dirichlet(np.full(spec.K, 0.6), size=2)
    n_all = spec.n_inds + spec.n_extra_inds
    rows = n_all * spec.ploidy * spec.nsamples
    dtype = np.uint16 if spec.n_donor_haps + n_all * spec.ploidy < 65535 else np.int32
    labels = np.empty((rows, spec.n_snps), dtype=dtype)
    for r in range(rows):
        labels[r] = _paint_one(rng, spec, cm, srcvec)
    return pos, rate, labels


def file_order(spec: SynthSpec):
    """Names of the individuals in file order and the 0-based file index of each recipient.
    Extras (non-recipients) are interleaved deterministically so that the 'skip' logic of
    the companion (reference C:418-425, C:553) is exercised."""
    n_all = spec.n_inds + spec.n_extra_inds
    names = [None] * n_all
    if spec.n_extra_inds == 0:
        slots = list(range(spec.n_inds))
    else:
        rng = np.random.default_rng(spec.seed + 99)
        slots = sorted(rng.choice(n_all, spec.n_inds, replace=False).tolist())
    for i, s in enumerate(slots):
        names[s] = spec.rec_labels[i]
    e = 0
    for j in range(n_all):
        if names[j] is None:
            e += 1
            names[j] = f"Other{e}"
    return names, slots


def _format_rows(mat: np.ndarray, width: int) -> np.ndarray:
    """Vectorised fixed-width decimal formatting: (R, L) ints -> (R, L*(width+1)) bytes
    (right-aligned, space separated; the reference tokenizer skips any whitespace, C:34)."""
    R, L = mat.shape
    out = np.full((R, L, width + 1), ord(" "), dtype=np.uint8)
    v = mat.astype(np.int64)
    first = np.ones_like(v, dtype=bool)
    for d in range(width):
        digit = v % 10
        nz = (v > 0) | first
        out[:, :, width - d][nz] = (ord("0") + digit[nz]).astype(np.uint8)
        v //= 10
        first[:] = False
    return out.reshape(R, L * (width + 1))


def write_samples_file(path: str, spec: SynthSpec, labels: np.ndarray, compresslevel: int = 1):
    names, _ = file_order(spec)
    width = len(str(int(labels.max())))
    header = (f"EM_iter = 0 (N_e = 0 / copy_prop = 0 / mutation = 0 / mutationGLOBAL = 0), nsamples = {spec.nsamples}, "
              f"N_e_start = 400.000000, region_size = 100.000000, haplotype dataset = {spec.name}.txt, "
              f"genmap dataset = {spec.name}.recomrates, pop-list dataset = popfile.txt, id-label dataset = individual.txt\n")
    opener = gzip.open if path.endswith(".gz") else open
    kw = {"compresslevel": compresslevel} if path.endswith(".gz") else {}
    with opener(path, "wb", **kw) as f:
        f.write(header.encode())
        r = 0
        for name in names:
            for p in range(spec.ploidy):
                f.write(f"HAP {p + 1} {name}\n".encode())
                block = _format_rows(labels[r:r + spec.nsamples], width)
                for s in range(spec.nsamples):
                    f.write(f"{s + 1}".encode())
                    f.write(block[s].tobytes())
                    f.write(b"\n")
                r += spec.nsamples


def write_recomb_file(path: str, pos: np.ndarray, rate: np.ndarray):
    with open(path, "w") as f:
        f.write("start.pos recom.rate.perbp\n")
        for p, r in zip(pos, rate):
            f.write(f"{int(p)} {r:.20f}\n")


@dataclass
class SynthData:
    spec: SynthSpec
    samples_list: str
    recom_list: str
    chrom: list          # list of dicts: pos, rate (as re-read from the text file!), labels
    donor_label_vec: np.ndarray
    rec_labels: list
    rec_file_index: list


def read_recomb_file(path: str):
    a = np.loadtxt(path, skiprows=1, dtype=np.float64, ndmin=2)
    return np.ascontiguousarray(a[:, 0]), np.ascontiguousarray(a[:, 1])


def write_dataset(spec: SynthSpec, outdir: str, gz: bool = True) -> SynthData:
    """Generate every chromosome, write the text inputs of the reference companion under
    ``outdir`` and return the packed arrays.  pos/rate are re-read from the text (what the
    companion sees after its decimal round trip)."""
    os.makedirs(outdir, exist_ok=True)
    chrom = []
    samp_paths, rec_paths = [], []
    for h in range(spec.n_chrom):
        pos, rate, labels = make_chromosome(spec, h)
        sp = os.path.join(outdir, f"{spec.name}_chr{h + 1}.samples.out" + (".gz" if gz else ""))
        rp = os.path.join(outdir, f"{spec.name}_chr{h + 1}.recomrates")
        write_samples_file(sp, spec, labels)
        write_recomb_file(rp, pos, rate)
        pos2, rate2 = read_recomb_file(rp)
        chrom.append({"pos": pos2, "rate": rate2, "labels": labels})
        samp_paths.append(sp)
        rec_paths.append(rp)
    sl = os.path.join(outdir, f"{spec.name}.samplefile.txt")
    rl = os.path.join(outdir, f"{spec.name}.recomfile.txt")
    with open(sl, "w") as f:
        f.write("\n".join(samp_paths) + "\n")
    with open(rl, "w") as f:
        f.write("\n".join(rec_paths) + "\n")
    _, slots = file_order(spec)
    return SynthData(spec, sl, rl, chrom, spec.donor_label_vec(), list(spec.rec_labels), slots)


def random_weights(K: int, S: int, seed: int = 5) -> np.ndarray:
    """A plausible ``weights.mat`` (S surrogates x K donors, columns... rows normalised like
    R:1186-1189 does: each surrogate's mixture over donors)."""
    rng = np.random.default_rng(seed)
    w = rng.dirichlet(np.full(S, 0.7), size=K).T  # S x K, each column sums to 1
    return np.ascontiguousarray(w)
