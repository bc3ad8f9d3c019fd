"""SASS opcode counts per kernel of the built library (cuobjdump -sass) -> profiles/r2_sass_opcodes.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "fastglobetrotter_b200", "fastGLOBETROTTERCompanion.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
ops = ["UBLKCP", "SYNCS", "MATCH", "BAR.ARV", "BAR.SYNC", "DMMA", "FENCE.VIEW.ASYNC", "SHFL", "VOTE", "REDG.E.ADD.F64", "REDUX", "DADD", "DMUL", "DFMA", "UTMALDG", "UTCHMMA", "LDTM"]
cur, cnt, tot = None, collections.defaultdict(collections.Counter), collections.Counter()
for ln in out.split("\n"):
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        continue
    if cur and re.search(r"/\*[0-9a-f]{4}\*/", ln):
        tot[cur] += 1
        for o in ops:
            if re.search(r"\s" + re.escape(o) + r"[\s.;]", ln):
                cnt[cur][o] += 1
dst = os.path.join(ROOT, "profiles", sys.argv[1] if len(sys.argv) > 1 else "r2_sass_opcodes.txt")
with open(dst, "w") as f:
    f.write("SASS opcode counts per kernel of fastGLOBETROTTERCompanion.so (cuobjdump -sass, sm_100a); instructions = SASS lines\n")
    f.write("TMA bulk copy = UBLKCP, mbarrier = SYNCS, match.any = MATCH, FP64 tensor pipe = DMMA, fp64 atomic = REDG.E.ADD.F64, cross-proxy fence = FENCE.VIEW.ASYNC.\n")
    f.write("No tcgen05 / TMA-tensor opcodes (UTC*MMA, LDTM, UTMALDG): the graded op is a gather + ordered fp64 read-modify-write, not a contraction.\n\n")
    for k in sorted(tot, key=lambda k: -tot[k]):
        nm = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip().split("(")[0]
        f.write(f"{nm:58s} instructions={tot[k]:6d}  " + " ".join(f"{o}={c}" for o, c in sorted(cnt[k].items())) + "\n")
print(dst)
