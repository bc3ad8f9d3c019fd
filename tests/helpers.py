"""Shared test helpers: locating libraries, tutorial-style input preparation (the part of
fastGLOBETROTTER.R that builds the arguments of the .C calls), libc srand()."""
from __future__ import annotations

import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REF_DIR = os.path.join(ROOT, "oracle", "_ref")
REF_LIB = os.path.join(REF_DIR, "fastGLOBETROTTERCompanion.so")
REF_HOOKED = os.path.join(REF_DIR, "fastGLOBETROTTERCompanion_hooked.so")
ORACLE_LIB = os.path.join(ROOT, "oracle", "liboracle.so")
GOLDEN = os.path.join(ROOT, "tests", "golden")
TUTORIAL = "/root/reference/tutorial"

_libc = ctypes.CDLL(None)


def srand(seed: int):
    """Seed the process-global libc rand() the reference draws from (C:136-137)."""
    _libc.srand(ctypes.c_uint(seed))


def libc_rand() -> int:
    return _libc.rand()


def have_ref() -> bool:
    return os.path.exists(REF_LIB) and os.path.exists(REF_HOOKED)


def id_file_inputs(id_path: str, donor_pops: list, target_pop: str, ploidy: int = 2):
    """fastGLOBETROTTER.R:805-828: donor.label.vec (one label per haplotype of the id file:
    1..K donor population, 0 target population, -9 otherwise) and recipient.label.vec."""
    rows = [ln.split() for ln in open(id_path) if ln.strip()]
    names = [r[0] for r in rows]
    pops = [r[1] for r in rows]
    incl = [r[2] != "0" for r in rows]
    lab = np.full(ploidy * len(rows), -9, dtype=np.int32)
    donors_temp = [p for p in donor_pops if p != target_pop]
    for i, dp in enumerate(donors_temp, start=1):
        for r in range(len(rows)):
            if pops[r] == dp and incl[r]:
                lab[ploidy * r: ploidy * (r + 1)] = i
    for r in range(len(rows)):
        if pops[r] == target_pop and incl[r]:
            lab[ploidy * r: ploidy * (r + 1)] = 0
    rec = [names[r] for r in range(len(rows)) if incl[r] and pops[r] == target_pop]
    return lab, rec


def parse_paramfile(path: str) -> dict:
    out = {}
    for ln in open(path):
        if ":" in ln:
            k, v = ln.split(":", 1)
            out[k.strip()] = v.strip()
    return out


def ind_id_matrix(n_inds: int, n_chrom: int) -> np.ndarray:
    """R:1059: matrix(rep(1:ninds-1, each=num.samplefiles), nrow=num.samplefiles) flattened
    column-major -> element [i*n_chrom + h] = i."""
    return np.repeat(np.arange(n_inds, dtype=np.int32), n_chrom)


def bootstrap_ids(n_inds: int, n_chrom: int, seed: int) -> np.ndarray:
    """R:2058-2059: per chromosome, resample individuals with replacement and sort."""
    rng = np.random.default_rng(seed)
    m = np.sort(rng.integers(0, n_inds, size=(n_chrom, n_inds)), axis=1)  # [h, i]
    return np.ascontiguousarray(m.T.ravel().astype(np.int32))  # [i*n_chrom + h]
