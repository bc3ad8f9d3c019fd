"""ctypes driver of the CPU oracle (oracle/companion_oracle.c).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs,
never by the product package."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "companion_oracle.c")
LIB = os.path.join(HERE, "liboracle.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build(force: bool = False) -> str:
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", LIB, SRC, "-lm"])
    return LIB


class Oracle:
    def __init__(self):
        self.lib = C.CDLL(build())
        L = self.lib
        L.orc_rand_sizeof.restype = C.c_size_t
        self._rng = C.create_string_buffer(L.orc_rand_sizeof())
        L.orc_rand.restype = C.c_int
        L.orc_data_read_packed.restype = C.c_long
        L.orc_data_read_null_packed.restype = C.c_long
        L.orc_chunk_one.restype = C.c_int
        self.srand(1)

    # --- glibc rand() emulator -------------------------------------------------
    def srand(self, seed: int):
        self.lib.orc_srand(self._rng, C.c_uint(seed))

    def rand(self) -> int:
        return self.lib.orc_rand(self._rng)

    def set_history(self, hist):
        h = np.ascontiguousarray(hist, dtype=np.uint32)
        assert h.size == 31
        self.lib.orc_rand_set_history(self._rng, h.ctypes.data_as(C.c_void_p))

    def get_history(self) -> np.ndarray:
        h = np.zeros(31, dtype=np.uint32)
        self.lib.orc_rand_get_history(self._rng, h.ctypes.data_as(C.c_void_p))
        return h

    # --- packed entry points ----------------------------------------------------
    @staticmethod
    def _chrom_arrays(chrom):
        n = len(chrom)
        nsites = np.array([c["pos"].size for c in chrom], dtype=np.int32)
        pos = [np.ascontiguousarray(c["pos"], dtype=np.float64) for c in chrom]
        rate = [np.ascontiguousarray(c["rate"], dtype=np.float64) for c in chrom]
        labs = [np.ascontiguousarray(c["labels"]) for c in chrom]
        lb = labs[0].dtype.itemsize
        assert lb in (2, 4) and all(l.dtype.itemsize == lb for l in labs)
        pp = (C.c_void_p * n)(*[p.ctypes.data for p in pos])
        rp = (C.c_void_p * n)(*[r.ctypes.data for r in rate])
        lp = (C.c_void_p * n)(*[l.ctypes.data for l in labs])
        return nsites, pp, rp, lp, lb, (pos, rate, labs)

    def data_read_packed(self, mode, ploidy, nsamples, chrom, ind_rows, bin_width, num_bins, grid_start, K,
                         donor_label_vec, weight_min, cutoff, S, temp_weights, want_raw=False):
        """Returns (means [S*S, G], raw or None, pairs)."""
        nsites, pp, rp, lp, lb, keep = self._chrom_arrays(chrom)
        ncho = len(chrom)
        ids = np.ascontiguousarray(ind_rows, dtype=np.int32).ravel()
        N = ids.size // ncho
        dl = np.ascontiguousarray(donor_label_vec, dtype=np.int32)
        tw = np.ascontiguousarray(temp_weights, dtype=np.float64).ravel()
        means = np.zeros(S * S * num_bins, dtype=np.float64)
        raw = None
        if want_raw:
            raw = np.zeros(((N if mode == 1 else N * ncho), K * K, num_bins), dtype=np.float64)
        pairs = self.lib.orc_data_read_packed(
            C.c_int(mode), C.c_int(ploidy), C.c_int(nsamples), C.c_int(ncho), C.c_int(N), ids.ctypes.data_as(_ip),
            nsites.ctypes.data_as(_ip), pp, rp, lp, C.c_int(lb), C.c_double(bin_width), C.c_int(num_bins),
            C.c_int(grid_start), C.c_int(K), dl.ctypes.data_as(_ip), C.c_double(weight_min), C.c_double(cutoff),
            C.c_int(S), tw.ctypes.data_as(_dp), self._rng, means.ctypes.data_as(_dp),
            raw.ctypes.data_as(_dp) if raw is not None else None)
        del keep
        return means.reshape(S * S, num_bins), raw, int(pairs)

    def data_read_null_packed(self, ploidy, nsamples, chrom, rec_rows, bin_width, num_bins, grid_start, K,
                              donor_label_vec, weight_min, cutoff, init=None):
        nsites, pp, rp, lp, lb, keep = self._chrom_arrays(chrom)
        rows = np.ascontiguousarray(rec_rows, dtype=np.int32)
        dl = np.ascontiguousarray(donor_label_vec, dtype=np.int32)
        # the reference adds into the caller's buffer (C:1020); `init` is that buffer's content on entry
        out = np.zeros(K * K * num_bins, dtype=np.float64) if init is None else \
            np.ascontiguousarray(init, dtype=np.float64).ravel().copy()
        evals = self.lib.orc_data_read_null_packed(
            C.c_int(ploidy), C.c_int(nsamples), C.c_int(len(chrom)), C.c_int(rows.size), rows.ctypes.data_as(_ip),
            nsites.ctypes.data_as(_ip), pp, rp, lp, C.c_int(lb), C.c_double(bin_width), C.c_int(num_bins),
            C.c_int(grid_start), C.c_int(K), dl.ctypes.data_as(_ip), C.c_double(weight_min), C.c_double(cutoff),
            out.ctypes.data_as(_dp))
        del keep
        return out.reshape(K * K, num_bins), int(evals)

    def mutiply_temp_weight(self, K, S, weights_mat_colmajor):
        w = np.ascontiguousarray(weights_mat_colmajor, dtype=np.float64).ravel()
        out = np.zeros(S * S * K * K, dtype=np.float64)
        self.lib.orc_mutiply_temp_weight(out.ctypes.data_as(_dp), w.ctypes.data_as(_dp), C.c_int(S), C.c_int(K))
        return out

    def chunk_one(self, labels_row, pos, rate, weight_min, cutoff, donor_label_vec):
        lab = np.ascontiguousarray(labels_row)
        L = lab.size
        p = np.ascontiguousarray(pos, dtype=np.float64)
        r = np.ascontiguousarray(rate, dtype=np.float64)
        dl = np.ascontiguousarray(donor_label_vec, dtype=np.int32)
        cap = L
        arrs = [np.zeros(cap, dtype=np.float64) for _ in range(4)]
        pop = np.zeros(cap, dtype=np.int32)
        n = self.lib.orc_chunk_one(lab.ctypes.data_as(C.c_void_p), C.c_int(lab.dtype.itemsize), C.c_int(L),
                                   p.ctypes.data_as(_dp), r.ctypes.data_as(_dp), C.c_double(weight_min),
                                   C.c_double(cutoff), dl.ctypes.data_as(_ip), C.c_int(cap),
                                   *[a.ctypes.data_as(_dp) for a in arrs], pop.ctypes.data_as(_ip))
        if n < 0:
            return n, None
        return n, {"pos": arrs[0][:n], "size": arrs[1][:n], "weight": arrs[2][:n], "end": arrs[3][:n], "pop": pop[:n]}


if __name__ == "__main__":
    print(build(force=True))
