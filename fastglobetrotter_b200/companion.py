"""Host-side mirror of the R -> C boundary of fastGLOBETROTTER (reference
fastGLOBETROTTER.R:199-232).

The reference loads ``fastGLOBETROTTERCompanion.so`` with ``dyn.load`` and calls
four symbols through ``.C()``.  ``.C`` semantics are plain C: every argument is a
pointer (scalars are length-1 arrays, strings are ``char**``), R owns and
pre-zeroes the output buffers and the C side adds into them.  R is not installed
in this image, so this module reproduces that marshalling with ctypes.  The same
class drives any library exporting the four symbols — the B200 library built from
``csrc/`` (the product) and, from the tests only, the compiled reference.

Function names follow the R wrappers: ``coancestry.curves`` ->
:meth:`Companion.coancestry_curves` (R:207-212), ``coancestry.curves.mode3``
(R:199-204), ``coancestry.curves.null`` (R:221-232), ``mutiply.temp.weight``
(R:214-219).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Sequence

import numpy as np

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
#: the drop-in library this repo builds (same file name the R script dyn.load()s, R:154)
DEFAULT_LIB = os.path.join(_PKG_DIR, "fastGLOBETROTTERCompanion.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_spp = C.POINTER(C.c_char_p)


def _ints(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32).ravel())


def _dbls(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).ravel())


def _strs(seq: Sequence[str]):
    arr = (C.c_char_p * len(seq))()
    arr[:] = [s.encode() for s in seq]
    return arr


class Companion:
    """ctypes binding of the four ``.C`` entry points of a companion library."""

    def __init__(self, path: str | None = None, mode: int = C.RTLD_LOCAL):
        self.path = path or DEFAULT_LIB
        if not os.path.exists(self.path):
            raise FileNotFoundError(
                f"{self.path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)")
        self.lib = C.CDLL(self.path, mode=mode)
        L = self.lib
        L.mutiply_temp_weight.argtypes = [_dp, _dp, _ip, _ip]
        L.mutiply_temp_weight.restype = None
        sig16 = [_ip, _ip, _spp, _spp, _dp, _ip, _ip, _ip, _ip, _ip, _spp, _dp, _dp, _ip, _dp, _dp]
        for name in ("data_read", "data_read_mode3"):
            f = getattr(L, name)
            f.argtypes = sig16
            f.restype = None
        L.data_read_null.argtypes = [_ip, _spp, _spp, _dp, _ip, _ip, _ip, _ip, _ip, _spp, _dp, _dp, _dp]
        L.data_read_null.restype = None

    # -- R:214-219 ---------------------------------------------------------
    def mutiply_temp_weight(self, unique_pops: int, num_surrogates: int, weights_mat) -> np.ndarray:
        """``weights_mat`` is the R matrix ``weights.mat`` (num_surrogates x unique_pops)
        flattened column-major as ``as.double`` does.  Returns the flat S^2*K^2 buffer
        exactly as C wrote it (row (m,n), column (k,l))."""
        w = _dbls(np.asarray(weights_mat, dtype=np.float64).ravel(order="F")) if np.ndim(weights_mat) == 2 else _dbls(weights_mat)
        out = np.zeros(num_surrogates * num_surrogates * unique_pops * unique_pops, dtype=np.float64)
        self.lib.mutiply_temp_weight(out.ctypes.data_as(_dp), w.ctypes.data_as(_dp),
                                     C.byref(C.c_int(num_surrogates)), C.byref(C.c_int(unique_pops)))
        return out

    def temp_weights_for_data_read(self, unique_pops: int, num_surrogates: int, weights_mat) -> np.ndarray:
        """The reshuffle R does between ``mutiply.temp.weight`` and ``coancestry.curves``
        (R:1195): ``matrix(x, byrow=TRUE, ncol=K*K)`` then ``as.double`` (column-major), so
        ``data_read`` indexes it as ``[(K*k+h)*S*S + S*m+n]`` (reference C:590)."""
        flat = self.mutiply_temp_weight(unique_pops, num_surrogates, weights_mat)
        mat = flat.reshape(num_surrogates * num_surrogates, unique_pops * unique_pops)
        return np.ascontiguousarray(mat.ravel(order="F"))

    # -- R:199-212 ---------------------------------------------------------
    def _curves(self, sym, ploidy, ind_id_vec, infile_samples, infile_recom, bin_width, num_bins, grid_start,
                num_donors, pop_label_vec, rec_labels, weightmin, gridweight_max_cutoff, temp_weights,
                num_surrogates) -> np.ndarray:
        ids = _ints(ind_id_vec)
        labels = _ints(pop_label_vec)
        tw = _dbls(temp_weights)
        means = np.zeros(num_surrogates * num_surrogates * num_bins, dtype=np.float64)
        samp, rec, recl = _strs([infile_samples]), _strs([infile_recom]), _strs(list(rec_labels))
        getattr(self.lib, sym)(
            C.byref(C.c_int(ploidy)), ids.ctypes.data_as(_ip), samp, rec,
            C.byref(C.c_double(bin_width)), C.byref(C.c_int(num_bins)), C.byref(C.c_int(grid_start)),
            C.byref(C.c_int(num_donors)), labels.ctypes.data_as(_ip), C.byref(C.c_int(len(rec_labels))), recl,
            C.byref(C.c_double(weightmin)), C.byref(C.c_double(gridweight_max_cutoff)),
            C.byref(C.c_int(num_surrogates)), tw.ctypes.data_as(_dp), means.ctypes.data_as(_dp))
        return means.reshape(num_surrogates * num_surrogates, num_bins)

    def coancestry_curves(self, *a, **k) -> np.ndarray:
        """``.C("data_read", ...)`` (R:207-212).  ``ind_id_vec`` is R's matrix
        [num_chrom x num_inds] flattened column-major, i.e. element ``[i*num_chrom + h]``."""
        return self._curves("data_read", *a, **k)

    def coancestry_curves_mode3(self, *a, **k) -> np.ndarray:
        """``.C("data_read_mode3", ...)`` (R:199-204)."""
        return self._curves("data_read_mode3", *a, **k)

    # -- R:221-232 ---------------------------------------------------------
    def coancestry_curves_null(self, ploidy, infile_samples, infile_recom, bin_width, num_bins, grid_start,
                               num_donors, pop_label_vec, rec_labels, weightmin, gridweight_max_cutoff,
                               target_index: int | None = None) -> np.ndarray:
        """``.C("data_read_null", ...)``; returns the K^2 x num_bins matrix.  If
        ``target_index`` (0-based position of the target population among the donor
        populations) is given, the target's rows and columns are dropped as R:225-230 does."""
        labels = _ints(pop_label_vec)
        out = np.zeros(num_donors * num_donors * num_bins, dtype=np.float64)
        samp, rec, recl = _strs([infile_samples]), _strs([infile_recom]), _strs(list(rec_labels))
        self.lib.data_read_null(
            C.byref(C.c_int(ploidy)), samp, rec, C.byref(C.c_double(bin_width)), C.byref(C.c_int(num_bins)),
            C.byref(C.c_int(grid_start)), C.byref(C.c_int(num_donors)), labels.ctypes.data_as(_ip),
            C.byref(C.c_int(len(rec_labels))), recl, C.byref(C.c_double(weightmin)),
            C.byref(C.c_double(gridweight_max_cutoff)), out.ctypes.data_as(_dp))
        mat = out.reshape(num_donors * num_donors, num_bins)
        if target_index is not None:
            K, t = num_donors, target_index
            drop = sorted(set(range(t, K * K, K)) | set(range(t * K, (t + 1) * K)))
            mat = np.delete(mat, drop, axis=0)
        return mat


def grid_start_for(curve_lower: float, curve_upper: float, bin_width: float) -> int:
    """R:1198-1202: 0-based index of ``curve.range[1]`` on the grid ``seq(0, upper, bw)``."""
    grid = np.arange(0, int(np.floor(curve_upper / bin_width + 1e-9)) + 1) * bin_width
    shift = np.abs(grid - curve_lower)
    return int(np.argmin(shift))


def num_bins_for(curve_lower: float, curve_upper: float, bin_width: float) -> int:
    """R:139-140: ``length(seq(lower, upper, bw))``."""
    return int(np.floor((curve_upper - curve_lower) / bin_width + 1e-10)) + 1


# ---------------------------------------------------------------------------------------------
# Packed boundary (include/fgt_companion.h part 2)
# ---------------------------------------------------------------------------------------------
class fgt_chrom_t(C.Structure):
    _fields_ = [("nsites", C.c_int32), ("pos", C.c_void_p), ("rate", C.c_void_p), ("labels", C.c_void_p),
                ("labels_on_device", C.c_int32), ("runs_on_device", C.c_int32), ("run_off", C.c_void_p),
                ("runs", C.c_void_p)]


#: numpy view of fgt_run_t: (start SNP, donor haplotype id)
RUN_DTYPE = np.dtype([("start", np.uint32), ("hap", np.uint32)])


class fgt_problem_t(C.Structure):
    _fields_ = [("mode", C.c_int32), ("ploidy", C.c_int32), ("nsamples", C.c_int32), ("num_chrom", C.c_int32),
                ("n_packed_inds", C.c_int32), ("label_bytes", C.c_int32), ("chrom", C.POINTER(fgt_chrom_t)),
                ("num_inds", C.c_int32), ("ind_rows", C.c_void_p), ("num_inds_total", C.c_int32),
                ("pair_offset", C.c_void_p), ("pairs_total", C.c_int64), ("binwidth", C.c_double),
                ("num_bins", C.c_int32), ("grid_start", C.c_int32), ("num_donors", C.c_int32),
                ("donor_label_vec", C.c_void_p), ("n_donor_labels", C.c_int32), ("weight_min", C.c_double),
                ("gridweight_max_cutoff", C.c_double), ("num_surrogates", C.c_int32), ("temp_weights", C.c_void_p)]


class fgt_stats_t(C.Structure):
    _fields_ = [("pairs", C.c_int64), ("accepted", C.c_int64), ("chunks", C.c_int64), ("kernel_launches", C.c_int32),
                ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64), ("ms_chunk", C.c_float), ("ms_rand", C.c_float),
                ("ms_accum", C.c_float), ("ms_project", C.c_float), ("ms_total", C.c_float),
                ("accum_launches", C.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


FGT_ERRORS = {0: "ok", -1: "no CUDA device (no CPU path)", -2: "CUDA error", -3: "missing donor (-9)",
              -4: "donor label out of range", -5: "bad argument", -6: "libc rand() state unsupported",
              -7: "mode 3 met a target-population chunk"}


class FgtError(RuntimeError):
    def __init__(self, rc):
        super().__init__(f"fgt error {rc}: {FGT_ERRORS.get(rc, '?')}")
        self.rc = rc


class PackedCompanion(Companion):
    """The B200 library including its packed entry points.  ``chrom`` arguments are lists of dicts
    with ``pos``/``rate`` (numpy float64) and ``labels`` -- a numpy array [rows, nsites] (uint16 or
    int32; host path) or a ``(device_ptr, dtype_itemsize)`` tuple (already resident in HBM)."""

    def __init__(self, path: str | None = None, encode: str = "rle"):
        """``encode``: how dense numpy label matrices handed to the packed calls reach the library -- "rle" (default:
        run-length encoded on the host first, the library's preferred painting format) or "dense" (as they are)."""
        super().__init__(path)
        assert encode in ("rle", "dense")
        self.encode = encode
        L = self.lib
        L.fgt_rle_encode.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_int64]
        L.fgt_rle_encode.restype = C.c_int
        L.fgt_pack_samples.argtypes = [C.c_char_p, C.c_int, C.c_char_p]
        L.fgt_pack_samples.restype = C.c_int64
        for nm in ("fgt_null_normalise", "fgt_getlhood_grid", "fgt_getmax", "fgt_bootstrap_means"):
            getattr(L, nm).restype = C.c_int
        L.fgt_dist_unique_id.restype = C.c_int
        L.fgt_dist_init.restype = C.c_int
        L.fgt_dist_finalize.restype = None
        L.fgt_data_read_packed.argtypes = [C.POINTER(fgt_problem_t), _dp, C.POINTER(fgt_stats_t)]
        L.fgt_data_read_packed.restype = C.c_int
        L.fgt_count_pairs_packed.argtypes = [C.POINTER(fgt_problem_t), C.POINTER(C.c_int64)]
        L.fgt_count_pairs_packed.restype = C.c_int
        L.fgt_data_read_null_packed.argtypes = [C.POINTER(fgt_problem_t), _ip, _dp, C.POINTER(fgt_stats_t)]
        L.fgt_data_read_null_packed.restype = C.c_int
        L.fgt_get_raw_counts.argtypes = [_dp, C.c_int64]
        L.fgt_get_raw_counts.restype = C.c_int64
        L.fgt_set_option.argtypes = [C.c_char_p, C.c_double]
        L.fgt_set_option.restype = C.c_int
        L.fgt_get_option.argtypes = [C.c_char_p]
        L.fgt_get_option.restype = C.c_double
        L.fgt_srand.argtypes = [C.c_uint]
        L.fgt_srand.restype = None
        L.fgt_last_stats.argtypes = [C.POINTER(fgt_stats_t)]
        L.fgt_last_stats.restype = None
        L.fgt_rand_fill.argtypes = [C.c_void_p, C.c_int64]
        L.fgt_rand_fill.restype = C.c_int
        L.fgt_chunk_one.restype = C.c_int
        L.fgt_version.restype = C.c_char_p

    # -- multi-process sharding (one process per GPU): NCCL communicator inside the library -----------------
    def dist_unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        rc = self.lib.fgt_dist_unique_id(buf)
        if rc:
            raise FgtError(rc)
        return buf.raw

    def dist_init(self, rank: int, world: int, uid: bytes):
        assert len(uid) == 128
        rc = self.lib.fgt_dist_init(C.c_int(rank), C.c_int(world), C.create_string_buffer(uid, 128))
        if rc:
            raise FgtError(rc)

    def dist_finalize(self):
        self.lib.fgt_dist_finalize()

    # -- callers of the path (include/fgt_companion.h part 3) -----------------------------------------------
    def null_normalise(self, temp_weights, null_mat, S, K, G, means=None):
        tw, nm = _dbls(temp_weights), _dbls(null_mat)
        m = _dbls(means).copy() if means is not None else None
        nc = np.zeros(S * S * G, dtype=np.float64)
        rc = self.lib.fgt_null_normalise(tw.ctypes.data_as(_dp), nm.ctypes.data_as(_dp), C.c_int(S), C.c_int(K), C.c_int(G),
                                         m.ctypes.data_as(_dp) if m is not None else None, nc.ctypes.data_as(_dp))
        if rc:
            raise FgtError(rc)
        return (m.reshape(S * S, G) if m is not None else None), nc.reshape(S * S, G)

    def getlhood_grid(self, means, times, admixtimes, popfracs, likpenalty=1.0):
        m = np.ascontiguousarray(means, dtype=np.float64)
        t, pf = _dbls(times), _dbls(popfracs)
        a = np.ascontiguousarray(admixtimes, dtype=np.float64)
        a = a.reshape(-1, 1) if a.ndim == 1 else a
        out = np.zeros(a.shape[0], dtype=np.float64)
        rc = self.lib.fgt_getlhood_grid(m.ctypes.data_as(_dp), C.c_int(m.shape[0]), C.c_int(m.shape[1]), t.ctypes.data_as(_dp),
                                        a.ctypes.data_as(_dp), C.c_int(a.shape[0]), C.c_int(a.shape[1]), pf.ctypes.data_as(_dp),
                                        C.c_double(likpenalty), out.ctypes.data_as(_dp))
        if rc:
            raise FgtError(rc)
        return out

    def getmax(self, means, times, nparam, popfracs):
        m = np.ascontiguousarray(means, dtype=np.float64)
        t, pf = _dbls(times), _dbls(popfracs)
        par = np.zeros(nparam, dtype=np.float64)
        val, pen = C.c_double(), C.c_double()
        rc = self.lib.fgt_getmax(m.ctypes.data_as(_dp), C.c_int(m.shape[0]), C.c_int(m.shape[1]), t.ctypes.data_as(_dp), C.c_int(nparam),
                                 pf.ctypes.data_as(_dp), par.ctypes.data_as(_dp), C.byref(val), C.byref(pen))
        if rc:
            raise FgtError(rc)
        return par, val.value, pen.value

    def bootstrap_means(self, ind_id_vecs, S, G):
        ids = np.ascontiguousarray(ind_id_vecs, dtype=np.int32)
        ids = ids.reshape(1, -1) if ids.ndim == 1 else ids
        out = np.zeros((ids.shape[0], S * S * G), dtype=np.float64)
        rc = self.lib.fgt_bootstrap_means(ids.ctypes.data_as(_ip), C.c_int(ids.shape[0]), out.ctypes.data_as(_dp))
        if rc:
            raise FgtError(rc)
        return out.reshape(ids.shape[0], S * S, G)

    def set_option(self, key: str, value: float):
        rc = self.lib.fgt_set_option(key.encode(), float(value))
        if rc:
            raise FgtError(rc)

    def get_option(self, key: str) -> float:
        return self.lib.fgt_get_option(key.encode())

    def srand(self, seed: int):
        self.lib.fgt_srand(C.c_uint(seed))

    def last_stats(self) -> dict:
        s = fgt_stats_t()
        self.lib.fgt_last_stats(C.byref(s))
        return s.as_dict()

    def rand_fill(self, n: int) -> np.ndarray:
        out = np.zeros(n, dtype=np.int32)
        rc = self.lib.fgt_rand_fill(out.ctypes.data, n)
        if rc:
            raise FgtError(rc)
        return out

    def rle_encode(self, labels):
        """dense [rows, nsites] uint16/int32 label matrix -> (run_off int64 [rows+1], runs RUN_DTYPE [n_runs])"""
        lab = np.ascontiguousarray(labels)
        assert lab.ndim == 2 and lab.dtype.itemsize in (2, 4)
        rows, L = lab.shape
        off = np.zeros(rows + 1, dtype=np.int64)
        n = self.lib.fgt_rle_encode(lab.ctypes.data, lab.dtype.itemsize, rows, L, off.ctypes.data, None, 0)
        if n < 0:
            raise FgtError(n)
        runs = np.zeros(n, dtype=RUN_DTYPE)
        n2 = self.lib.fgt_rle_encode(lab.ctypes.data, lab.dtype.itemsize, rows, L, off.ctypes.data, runs.ctypes.data, n)
        assert n2 == n
        return off, runs

    def pack_samples(self, samples_path: str, nsites: int, out_path: str) -> int:
        n = self.lib.fgt_pack_samples(samples_path.encode(), nsites, out_path.encode())
        if n < 0:
            raise FgtError(int(n))
        return int(n)

    def _problem(self, mode, ploidy, nsamples, chrom, ind_rows, bin_width, num_bins, grid_start, K, donor_label_vec,
                 weight_min, cutoff, S, temp_weights, n_packed_inds=None, num_inds_total=0, pair_offset=None,
                 pairs_total=0):
        keep = []
        n = len(chrom)
        carr = (fgt_chrom_t * n)()
        lb = None
        rows = None
        for h, c in enumerate(chrom):
            pos = np.ascontiguousarray(c["pos"], dtype=np.float64)
            rate = np.ascontiguousarray(c["rate"], dtype=np.float64)
            keep += [pos, rate]
            carr[h].nsites = pos.size
            carr[h].pos = pos.ctypes.data
            carr[h].rate = rate.ctypes.data
            lab = c.get("labels")
            runs = c.get("runs")
            if runs is None and lab is not None and not isinstance(lab, tuple) and self.encode == "rle":
                runs = self.rle_encode(lab)
            if runs is not None:
                # (run_off int64 [rows+1], runs): runs a RUN_DTYPE array (host) or a device pointer (int)
                off, rr = runs
                off = np.ascontiguousarray(off, dtype=np.int64)
                keep.append(off)
                carr[h].run_off = off.ctypes.data
                if isinstance(rr, (int, np.integer)):
                    carr[h].runs = int(rr)
                    carr[h].runs_on_device = 1
                else:
                    rr = np.ascontiguousarray(rr)
                    assert rr.dtype.itemsize == 8 or (rr.ndim == 2 and rr.shape[1] == 2 and rr.dtype.itemsize == 4)
                    keep.append(rr)
                    carr[h].runs = rr.ctypes.data
                    carr[h].runs_on_device = 0
                this_lb, this_rows = 4, off.size - 1
            elif isinstance(lab, tuple):
                ptr, item, nrows = lab
                carr[h].labels = ptr
                carr[h].labels_on_device = 1
                this_lb, this_rows = item, nrows
            else:
                lab = np.ascontiguousarray(lab)
                keep.append(lab)
                carr[h].labels = lab.ctypes.data
                carr[h].labels_on_device = 0
                this_lb, this_rows = lab.dtype.itemsize, lab.shape[0]
            if runs is None:
                assert lb in (None, this_lb)
                lb = this_lb
            rows = this_rows
        ids = np.ascontiguousarray(ind_rows, dtype=np.int32).ravel()
        dl = np.ascontiguousarray(donor_label_vec, dtype=np.int32)
        keep += [ids, dl, carr]
        pb = fgt_problem_t()
        pb.mode, pb.ploidy, pb.nsamples, pb.num_chrom = mode, ploidy, nsamples, n
        pb.n_packed_inds = n_packed_inds if n_packed_inds is not None else rows // (ploidy * nsamples)
        pb.label_bytes = lb if lb is not None else 4
        pb.chrom = carr
        pb.num_inds = ids.size // n
        pb.ind_rows = ids.ctypes.data
        pb.num_inds_total = num_inds_total
        if pair_offset is not None:
            po = np.ascontiguousarray(pair_offset, dtype=np.int64)
            keep.append(po)
            pb.pair_offset = po.ctypes.data
            pb.pairs_total = int(pairs_total)
        pb.binwidth, pb.num_bins, pb.grid_start, pb.num_donors = bin_width, num_bins, grid_start, K
        pb.donor_label_vec = dl.ctypes.data
        pb.n_donor_labels = dl.size
        pb.weight_min, pb.gridweight_max_cutoff, pb.num_surrogates = weight_min, cutoff, S
        if temp_weights is not None:
            tw = np.ascontiguousarray(temp_weights, dtype=np.float64).ravel()
            keep.append(tw)
            pb.temp_weights = tw.ctypes.data
        return pb, keep

    def data_read_packed(self, mode, ploidy, nsamples, chrom, ind_rows, bin_width, num_bins, grid_start, K,
                         donor_label_vec, weight_min, cutoff, S, temp_weights, want_raw=False, means_init=None, **kw):
        """Returns (means [S*S, G], raw or None, stats dict)."""
        pb, keep = self._problem(mode, ploidy, nsamples, chrom, ind_rows, bin_width, num_bins, grid_start, K,
                                 donor_label_vec, weight_min, cutoff, S, temp_weights, **kw)
        means = np.zeros(S * S * num_bins, dtype=np.float64) if means_init is None else \
            np.ascontiguousarray(means_init, dtype=np.float64).ravel().copy()
        st = fgt_stats_t()
        self.set_option("keep_raw", 1 if want_raw else 0)
        rc = self.lib.fgt_data_read_packed(C.byref(pb), means.ctypes.data_as(_dp), C.byref(st))
        del keep
        if rc:
            raise FgtError(rc)
        raw = self.get_raw_counts(K, num_bins) if want_raw else None
        return means.reshape(S * S, num_bins), raw, st.as_dict()

    def count_pairs_packed(self, mode, ploidy, nsamples, chrom, ind_rows, bin_width, num_bins, grid_start, K,
                           donor_label_vec, weight_min, cutoff, **kw) -> np.ndarray:
        pb, keep = self._problem(mode, ploidy, nsamples, chrom, ind_rows, bin_width, num_bins, grid_start, K,
                                 donor_label_vec, weight_min, cutoff, 1, None, **kw)
        out = np.zeros(pb.num_chrom * pb.num_inds, dtype=np.int64)
        rc = self.lib.fgt_count_pairs_packed(C.byref(pb), out.ctypes.data_as(C.POINTER(C.c_int64)))
        del keep
        if rc:
            raise FgtError(rc)
        return out.reshape(pb.num_chrom, pb.num_inds)

    def get_raw_counts(self, K, num_bins) -> np.ndarray:
        n = self.lib.fgt_get_raw_counts(None, 0)
        out = np.zeros(max(n, 0), dtype=np.float64)
        if n > 0:
            self.lib.fgt_get_raw_counts(out.ctypes.data_as(_dp), n)
        return out.reshape(-1, K * K, num_bins)

    def data_read_null_packed(self, ploidy, nsamples, chrom, rec_rows, bin_width, num_bins, grid_start, K,
                              donor_label_vec, weight_min, cutoff, init=None):
        rows = np.ascontiguousarray(rec_rows, dtype=np.int32)
        pb, keep = self._problem(1, ploidy, nsamples, chrom, np.zeros(len(chrom) * rows.size, dtype=np.int32), bin_width,
                                 num_bins, grid_start, K, donor_label_vec, weight_min, cutoff, 1, None)
        out = np.zeros(K * K * num_bins, dtype=np.float64) if init is None else \
            np.ascontiguousarray(init, dtype=np.float64).ravel().copy()
        st = fgt_stats_t()
        rc = self.lib.fgt_data_read_null_packed(C.byref(pb), rows.ctypes.data_as(_ip), out.ctypes.data_as(_dp), C.byref(st))
        del keep
        if rc:
            raise FgtError(rc)
        return out.reshape(K * K, num_bins), st.as_dict()

    def chunk_one(self, labels_row, pos, rate, weight_min, cutoff, donor_label_vec):
        lab = np.ascontiguousarray(labels_row)
        L = lab.size
        p = np.ascontiguousarray(pos, dtype=np.float64)
        r = np.ascontiguousarray(rate, dtype=np.float64)
        dl = np.ascontiguousarray(donor_label_vec, dtype=np.int32)
        arrs = [np.zeros(L, dtype=np.float64) for _ in range(4)]
        pop = np.zeros(L, dtype=np.int32)
        n = self.lib.fgt_chunk_one(C.c_void_p(lab.ctypes.data), C.c_int(lab.dtype.itemsize), C.c_int(L),
                                   p.ctypes.data_as(_dp), r.ctypes.data_as(_dp), C.c_double(weight_min),
                                   C.c_double(cutoff), dl.ctypes.data_as(_ip), C.c_int(dl.size),
                                   *[a.ctypes.data_as(_dp) for a in arrs], pop.ctypes.data_as(_ip))
        if n < 0:
            return n, None
        return n, {"pos": arrs[0][:n], "size": arrs[1][:n], "weight": arrs[2][:n], "end": arrs[3][:n], "pop": pop[:n]}
