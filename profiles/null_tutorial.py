"""data_read_null on the reference's tutorial (text files under tests/golden/tutorial) for each null_impl: time of the cached call,
check against the golden NULL tensor of the compiled reference."""
import os, sys, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import golden_util as gu
from helpers import GOLDEN, id_file_inputs, parse_paramfile
from fastglobetrotter_b200 import build as fbuild
from fastglobetrotter_b200.companion import PackedCompanion
lib = PackedCompanion(fbuild.build())
gold = gu.load("tutorial_full.npz")
tdir = os.path.join(GOLDEN, "tutorial")
prm = parse_paramfile(os.path.join(tdir, "paramfile.txt"))
dl, rec = id_file_inputs(os.path.join(tdir, "individual.txt"), prm["copyvector.popnames"].split(), prm["target.popname"])
K, bw, G, gs, stride = int(gold["K"]), float(gold["bw"]), int(gold["G"]), int(gold["gs"]), int(gold["stride"])
os.chdir(GOLDEN)
if os.environ.get("DEVICES"): lib.set_option("devices", int(os.environ["DEVICES"]))   # one process, several GPUs
for impl in [int(x) for x in os.environ.get("IMPLS", "1,2,0").split(",")]:
    lib.set_option("null_impl", impl)
    for rep in range(3):
        lib.set_option("verbose", 1 if (rep == 2 and os.environ.get("VERBOSE")) else 0)
        t0 = time.perf_counter()
        nul = lib.coancestry_curves_null(2, "tutorial/samplefile.txt", "tutorial/recomfile.txt", bw, G, gs, K, dl, rec, 0.0, 1.0)
        dt = time.perf_counter() - t0
    lib.set_option("verbose", 0)
    st = lib.last_stats()
    gu.assert_matches(gold, "null", nul, stride)
    print(f"null_impl={impl}: call {dt*1e3:.1f} ms, kernels {st['ms_accum']:.1f} ms, {st['pairs']/1e9:.2f} G evaluations, {st['pairs']/dt/1e9:.1f} G/s, accepted {st['accepted']/st['pairs']:.3f}, == reference", flush=True)
if os.environ.get("TASKS"):
    lib.set_option("null_impl", 2)
    for nt in [int(x) for x in os.environ["TASKS"].split(",")]:
        lib.set_option("null_tasks", nt)
        for rep in range(3):
            t0 = time.perf_counter()
            nul = lib.coancestry_curves_null(2, "tutorial/samplefile.txt", "tutorial/recomfile.txt", bw, G, gs, K, dl, rec, 0.0, 1.0)
            dt = time.perf_counter() - t0
        print(f"null_tasks={nt}: call {dt*1e3:.1f} ms, kernels {lib.last_stats()['ms_accum']:.1f} ms", flush=True)
