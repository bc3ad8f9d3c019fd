"""CPU tests of the product library's host side: the C-ABI library loads and exports every symbol
declared in include/fgt_companion.h; compute entry points fail loudly without a GPU (no CPU
fallback); the jump-ahead arithmetic and the multithreaded text/gz layer are correct."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from helpers import ROOT, libc_rand, srand
from fastglobetrotter_b200 import synth


def declared_functions():
    hdr = open(os.path.join(ROOT, "include", "fgt_companion.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*;", hdr)
    return sorted(set(n for n in names if n not in ("defined",)))


def test_library_exports_every_declared_symbol(lib):
    names = declared_functions()
    for must in ("data_read", "data_read_mode3", "data_read_null", "mutiply_temp_weight", "fgt_data_read_packed",
                 "fgt_data_read_null_packed", "fgt_count_pairs_packed", "fgt_parse_probe"):
        assert must in names
    for n in names:
        assert hasattr(lib.lib, n), f"{n} declared in include/fgt_companion.h but not exported"
    # the R-facing symbols are plain C symbols (dyn.load + .C need unmangled names)
    out = subprocess.run(["nm", "-D", "--defined-only", lib.path], capture_output=True, text=True).stdout
    for n in ("data_read", "data_read_mode3", "data_read_null", "mutiply_temp_weight"):
        assert re.search(rf"\sT {n}$", out, flags=re.M), n


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from fastglobetrotter_b200.companion import FgtError
    with pytest.raises(FgtError) as e:
        lib.rand_fill(16)
    assert e.value.rc == -1
    spec = synth.SynthSpec(n_chrom=1, n_inds=1, n_snps=200, K=2, nsamples=1, rate=5e-6)
    pos, rate, labels = synth.make_chromosome(spec, 0)
    with pytest.raises(FgtError):
        lib.data_read_packed(1, 2, 1, [{"pos": pos, "rate": rate, "labels": labels}], [0], 0.1, 20, 10, 2, spec.donor_label_vec(),
                             0.0, 1.0, 1, np.ones(4))
    # the .C entry points end the process like the reference does on fatal errors (printf + exit(1))
    code = ("import sys; sys.path.insert(0, %r); from fastglobetrotter_b200.companion import Companion; import numpy as np;"
            "Companion().mutiply_temp_weight(2, 1, np.ones((1, 2)))") % ROOT
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert p.returncode == 1 and "no CUDA device" in p.stdout


def test_jump_ahead_matches_stepping(lib, oracle):
    hist = np.zeros(31, dtype=np.uint32)
    for seed, n in ((1, 0), (1, 1), (7, 30), (7, 31), (12345, 1984), (99, 1_000_003), (5, 3_000_000_019)):
        assert lib.lib.fgt_rand_host_jump(C.c_uint(seed), C.c_uint64(n), hist.ctypes.data_as(C.c_void_p)) == 0
        if n <= 2_000_000:
            oracle.srand(seed)
            for _ in range(n):
                oracle.rand()
            assert np.array_equal(oracle.get_history(), hist), (seed, n)
        # continuing from the jumped state reproduces libc
        oracle.set_history(hist)
        if n <= 2_000_000:
            srand(seed)
            for _ in range(n):
                libc_rand()
            assert [oracle.rand() for _ in range(50)] == [libc_rand() for _ in range(50)]
    # composition: jump(a+b) == jump(b) after jump(a), checked through two different splits of a big offset
    big = 123_456_789_012
    lib.lib.fgt_rand_host_jump(C.c_uint(3), C.c_uint64(big), hist.ctypes.data_as(C.c_void_p))
    oracle.set_history(hist)
    a = [oracle.rand() for _ in range(10)]
    lib.lib.fgt_rand_host_jump(C.c_uint(3), C.c_uint64(big + 4), hist.ctypes.data_as(C.c_void_p))
    oracle.set_history(hist)
    assert [oracle.rand() for _ in range(6)] == a[4:]


def _probe(lib, data, labels_wanted=True):
    spec = data.spec
    n = spec.n_inds
    recl = (C.c_char_p * n)(*[s.encode() for s in data.rec_labels])
    num_chrom, nsamples, lb = C.c_int32(), C.c_int32(), C.c_int32()
    nsites = (C.c_int32 * 64)()
    rec_index = (C.c_int32 * n)()
    sums = (C.c_uint64 * 64)()
    rows = n * spec.ploidy * spec.nsamples
    bufs = [np.zeros((rows, spec.n_snps), dtype=np.uint16) for _ in range(spec.n_chrom)]
    ptrs = (C.c_void_p * spec.n_chrom)(*[b.ctypes.data for b in bufs])
    pos = np.zeros(spec.n_chrom * spec.n_snps)
    rate = np.zeros(spec.n_chrom * spec.n_snps)
    lib.lib.fgt_parse_probe.restype = C.c_int
    rc = lib.lib.fgt_parse_probe(data.samples_list.encode(), data.recom_list.encode(), C.c_int(spec.ploidy), C.c_int(n), recl,
                                 C.byref(num_chrom), C.byref(nsamples), nsites, rec_index, C.byref(lb), sums,
                                 ptrs if labels_wanted else None, pos.ctypes.data_as(C.c_void_p), rate.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return num_chrom.value, nsamples.value, list(nsites)[:num_chrom.value], list(rec_index), lb.value, bufs, pos, rate


@pytest.mark.parametrize("gz", [True, False])
def test_text_layer_decodes_what_was_written(lib, tmp_path, gz):
    spec = synth.SynthSpec(n_chrom=3, n_inds=4, n_snps=1500, K=5, nsamples=3, n_extra_inds=3, rate=6e-7, name="parse")
    data = synth.write_dataset(spec, str(tmp_path), gz=gz)
    nchrom, nsamples, nsites, rec_index, lb, bufs, pos, rate = _probe(lib, data)
    assert (nchrom, nsamples, lb) == (3, 3, 2)
    assert nsites == [1500] * 3
    assert rec_index == data.rec_file_index
    per = spec.ploidy * spec.nsamples
    for h in range(3):
        rows = np.concatenate([np.arange(s * per, (s + 1) * per) for s in data.rec_file_index])
        assert np.array_equal(bufs[h], data.chrom[h]["labels"][rows])
        assert np.array_equal(pos[h * 1500:(h + 1) * 1500], data.chrom[h]["pos"])
        assert np.array_equal(rate[h * 1500:(h + 1) * 1500], data.chrom[h]["rate"])
    # second call is served from the per-process cache and must give the same answer
    again = _probe(lib, data)
    assert np.array_equal(again[5][1], bufs[1])


def test_list_file_semantics_follow_the_reference(lib, tmp_path):
    """an unterminated last line of a list file is dropped (the reference's feof loop, C:203-208), and the
    number of SNPs is the number of newline-terminated map rows (C:357-362)."""
    spec = synth.SynthSpec(n_chrom=2, n_inds=2, n_snps=400, K=3, nsamples=2, rate=2e-6, name="lst")
    data = synth.write_dataset(spec, str(tmp_path))
    sl = open(data.samples_list).read().rstrip("\n")
    rl = open(data.recom_list).read().rstrip("\n")
    open(data.samples_list, "w").write(sl)      # no trailing newline -> second chromosome dropped
    open(data.recom_list, "w").write(rl)
    nchrom, *_ = _probe(lib, data, labels_wanted=False)
    assert nchrom == 1


def test_fatal_input_errors_exit_like_the_reference(lib, tmp_path):
    spec = synth.SynthSpec(n_chrom=1, n_inds=2, n_snps=300, K=3, nsamples=2, rate=2e-6, name="bad")
    data = synth.write_dataset(spec, str(tmp_path))
    base = ("import sys, ctypes as C; sys.path.insert(0, %r); from fastglobetrotter_b200.companion import PackedCompanion;"
            "l = PackedCompanion(); l.lib.fgt_parse_probe.restype = C.c_int;"
            "rec = (C.c_char_p * 2)(%s); n = C.c_int32();"
            "l.lib.fgt_parse_probe(%r.encode(), %r.encode(), C.c_int(2), C.c_int(2), rec, C.byref(n), None, None, None, None, None, None, None, None)")
    ok = base % (ROOT, "b'Target1', b'Target2'", data.samples_list, data.recom_list)
    assert subprocess.run([sys.executable, "-c", ok], capture_output=True, text=True).returncode == 0
    missing = base % (ROOT, "b'Target1', b'Nobody'", data.samples_list, data.recom_list)
    p = subprocess.run([sys.executable, "-c", missing], capture_output=True, text=True)
    assert p.returncode == 1 and "recipient haplotypes" in p.stdout
    nofile = base % (ROOT, "b'Target1', b'Target2'", "/nonexistent/list.txt", data.recom_list)
    p = subprocess.run([sys.executable, "-c", nofile], capture_output=True, text=True)
    assert p.returncode == 1 and "error opening" in p.stdout


def test_on_disk_pack_cache_round_trip(lib, tmp_path, monkeypatch):
    """FGT_PACK_CACHE_DIR: decoded paintings are written as binary files and a later reader (no in-process cache,
    even with the text gone stale in memory) gets the identical arrays without inflating; a changed source file,
    a different recipient set or a damaged cache file fall back to parsing the text."""
    spec = synth.SynthSpec(n_chrom=2, n_inds=3, n_snps=900, K=4, nsamples=3, n_extra_inds=2, rate=6e-7, name="pack")
    data = synth.write_dataset(spec, str(tmp_path / "data"))
    cache = tmp_path / "cache"
    cache.mkdir()
    monkeypatch.setenv("FGT_PACK_CACHE_DIR", str(cache))
    lib.set_option("cache", 0)                       # take the per-process cache out of the picture
    try:
        first = _probe(lib, data)
        packs = sorted(p.name for p in cache.iterdir())
        assert sum(n.endswith(".fgtpack") for n in packs) == 2 and sum(n.endswith(".fgtmeta") for n in packs) == 1
        assert not any(".tmp" in n for n in packs)
        # served from disk: make the text unreadable as paintings but keep size and mtime, so only the cache can answer
        import gzip
        sfile = open(data.samples_list).read().split()[0]
        st = os.stat(sfile)
        raw = open(sfile, "rb").read()
        open(sfile, "wb").write(b"\0" * len(raw))
        os.utime(sfile, ns=(st.st_atime_ns, st.st_mtime_ns))
        second = _probe(lib, data)
        assert second[:5] == first[:5]
        for a, b in zip(first[5], second[5]):
            assert np.array_equal(a, b)
        # restore the text with a new mtime: the stale cache entry is ignored and a fresh one is written
        open(sfile, "wb").write(raw)
        os.utime(sfile, ns=(st.st_atime_ns, st.st_mtime_ns + 5_000_000_000))
        third = _probe(lib, data)
        for a, b in zip(first[5], third[5]):
            assert np.array_equal(a, b)
        assert sum(p.name.endswith(".fgtpack") for p in cache.iterdir()) == 3
        # a truncated cache file is rejected
        for p in cache.iterdir():
            if p.name.endswith(".fgtpack"):
                b = p.read_bytes()
                p.write_bytes(b[: len(b) // 2])
        fourth = _probe(lib, data)
        for a, b in zip(first[5], fourth[5]):
            assert np.array_equal(a, b)
    finally:
        lib.set_option("cache", 1)


def test_rle_encode_and_interchange_file_round_trip(lib, tmp_path):
    """fgt_rle_encode (dense matrix -> runs) and the self-contained chunk-level file format: a text samples file is
    converted to *.fgtrle once, and a samples list naming the .fgtrle file decodes to the very same paintings."""
    spec = synth.SynthSpec(n_chrom=2, n_inds=4, n_snps=1200, K=5, nsamples=3, n_extra_inds=2, rate=6e-7, name="rle")
    data = synth.write_dataset(spec, str(tmp_path))
    lab = data.chrom[0]["labels"]
    off, runs = lib.rle_encode(lab)
    assert off[0] == 0 and off[-1] == runs.size and np.all(np.diff(off) >= 1)
    back = np.zeros_like(lab)
    for r in range(lab.shape[0]):
        rr = runs[off[r]:off[r + 1]]
        assert rr["start"][0] == 0 and np.all(np.diff(rr["start"].astype(np.int64)) > 0) and np.all(rr["hap"][1:] != rr["hap"][:-1])
        ends = np.append(rr["start"][1:], lab.shape[1])
        for s, e, hp in zip(rr["start"], ends, rr["hap"]):
            back[r, s:e] = hp
    assert np.array_equal(back, lab)
    text = _probe(lib, data)
    sfiles = open(data.samples_list).read().split()
    packed = []
    for h, sf in enumerate(sfiles):
        out = str(tmp_path / f"chr{h + 1}.fgtrle")
        n = lib.pack_samples(sf, spec.n_snps, out)
        assert n > 0 and os.path.getsize(out) == 48 + sum(len(x) + 1 for x in synth.file_order(spec)[0]) * spec.ploidy + 8 * (len(synth.file_order(spec)[0]) * spec.ploidy * spec.nsamples + 1) + 8 * n
        packed.append(out)
    open(data.samples_list, "w").write("\n".join(packed) + "\n")
    binary = _probe(lib, data)
    assert binary[:5] == text[:5]
    for a, b in zip(text[5], binary[5]):
        assert np.array_equal(a, b)


def test_truncated_gzip_is_fatal_and_cache_is_bounded(lib, tmp_path):
    """a samples file cut in the middle of its gzip stream ends the process with a message (zlib's error is not mistaken
    for a clean end of file); the per-process cache of decoded files obeys its byte budget ("cache_mb")."""
    spec = synth.SynthSpec(n_chrom=3, n_inds=3, n_snps=2000, K=4, nsamples=3, rate=6e-7, name="trunc")
    data = synth.write_dataset(spec, str(tmp_path))
    ok = _probe(lib, data)
    lib.set_option("cache_mb", 0.02)                       # ~20 KB: less than one decoded chromosome
    try:
        again = _probe(lib, data)                          # entries are evicted as they come in; results unchanged
        for a, b in zip(ok[5], again[5]):
            assert np.array_equal(a, b)
    finally:
        lib.set_option("cache_mb", 8192)
    sfile = open(data.samples_list).read().split()[1]
    raw = open(sfile, "rb").read()
    open(sfile, "wb").write(raw[: len(raw) // 2])
    code = ("import sys, ctypes as C; sys.path.insert(0, %r); from fastglobetrotter_b200.companion import PackedCompanion;"
            "l = PackedCompanion(); l.lib.fgt_parse_probe.restype = C.c_int;"
            "rec = (C.c_char_p * 3)(b'Target1', b'Target2', b'Target3'); n = C.c_int32();"
            "l.lib.fgt_parse_probe(%r.encode(), %r.encode(), C.c_int(2), C.c_int(3), rec, C.byref(n), None, None, None, None, None, None, None, None)"
            ) % (ROOT, data.samples_list, data.recom_list)
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert p.returncode == 1 and "Something wrong with" in p.stdout


def test_option_setters_for_R_and_environment(lib):
    """fgt_set_option_R is the .C flavour (every argument a pointer); FGT_DEVICES presets the "devices" option at first use."""
    key = (C.c_char_p * 1)(b"cache_mb")
    val = C.c_double(123.0)
    lib.lib.fgt_set_option_R(key, C.byref(val))
    assert lib.get_option("cache_mb") == 123.0
    lib.set_option("cache_mb", 8192)
    code = ("import sys; sys.path.insert(0, %r); from fastglobetrotter_b200.companion import PackedCompanion;"
            "print(int(PackedCompanion().get_option('devices')))") % ROOT
    env = dict(os.environ, FGT_DEVICES="4")
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env)
    assert p.returncode == 0 and p.stdout.strip() == "4"
