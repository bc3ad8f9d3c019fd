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
