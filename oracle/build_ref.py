#!/usr/bin/env python3
"""Build recipe for the *real* reference companion (TEST INFRASTRUCTURE ONLY).

Compiles /root/reference/fastGLOBETROTTERCompanion.c where it lies into
oracle/_ref/ (git-ignored, but it does travel to the GPU box with gpurun):

  oracle/_ref/fastGLOBETROTTERCompanion.so   pristine, `gcc -O2 -fPIC -shared ... -lz -lm`
                                             (the reference's own build line is
                                             `R CMD SHLIB ... -lz -O2`, README.md:8)
  oracle/_ref/fastGLOBETROTTERCompanion_hooked.so
                                             the same translation unit with three
                                             dump hooks spliced in at build time
                                             (never written to disk as source):
        - before the projection loop of data_read (reference C:578): copy
          total_counts_mat_all[n][k][i] into REF_DUMP_counts (if non-NULL)
        - at the end of tabulate_chunks / tabulate_chunks_mode3: REF_DUMP_pairs += count_ans
        - before the projection in tabulate_chunks_mode3 (reference C:1190): copy the
          per-(chromosome, individual) total_counts_mat into REF_DUMP_counts3 at slot
          REF_DUMP_calls3++ (if non-NULL)

No reference source is copied into the repository: the patched text only ever
exists in a pipe to gcc.  If /root/reference is absent (GPU box) this script is
a no-op and the prebuilt .so files are used.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/fastGLOBETROTTERCompanion.c"
OUT = os.path.join(HERE, "_ref")
PRISTINE = os.path.join(OUT, "fastGLOBETROTTERCompanion.so")
HOOKED = os.path.join(OUT, "fastGLOBETROTTERCompanion_hooked.so")

GLOBALS = """
double *REF_DUMP_counts = 0;   /* [N][K*K][G] */
double *REF_DUMP_counts3 = 0;  /* [calls][K*K][G] */
long REF_DUMP_calls3 = 0;
long REF_DUMP_pairs = 0;
"""

HOOK_MODE1 = """
if (REF_DUMP_counts) { long __o = 0; for (int __n=0; __n<*num_inds_rec; __n++) for (int __k=0; __k<(*num_donors)*(*num_donors); __k++) for (int __i=0; __i<*num_bins; __i++) REF_DUMP_counts[__o++] = total_counts_mat_all[__n][__k][__i]; }
"""
HOOK_PAIRS = "REF_DUMP_pairs += count_ans;\n"
HOOK_MODE3 = """
if (REF_DUMP_counts3) { long __o = REF_DUMP_calls3*(long)(*num_donors)*(*num_donors)*(*num_bins); for (int __k=0; __k<(*num_donors)*(*num_donors); __k++) for (int __i=0; __i<*num_bins; __i++) REF_DUMP_counts3[__o++] = total_counts_mat[__k][__i]; }
REF_DUMP_calls3++;
"""


def patched_source() -> str:
    lines = open(REF_SRC).read().split("\n")
    out = []
    seen = {"globals": 0, "mode1": 0, "pairs": 0, "mode3": 0}
    in_tab = None
    for ln in lines:
        s = ln.strip()
        if s.startswith("void mutiply_temp_weight") and not seen["globals"]:
            out.append(GLOBALS)
            seen["globals"] += 1
        if s.startswith("void tabulate_chunks_mode3("):
            in_tab = "m3"
        elif s.startswith("void tabulate_chunks("):
            in_tab = "m1"
        elif s.startswith("void data_read"):
            in_tab = None
        # mode-1 raw tensor: immediately before the projection loop
        if s.startswith("for (int ind_num=0; ind_num<*num_inds_rec;ind_num++)"):
            out.append(HOOK_MODE1)
            seen["mode1"] += 1
        # pair counters: right before the scratch frees at the end of both tabulate functions
        if in_tab and s.startswith("free(numchunckinbin);"):
            out.append(HOOK_PAIRS)
            seen["pairs"] += 1
            if in_tab == "m3":
                out.append(HOOK_MODE3)
                seen["mode3"] += 1
        out.append(ln)
    assert seen == {"globals": 1, "mode1": 1, "pairs": 2, "mode3": 1}, seen
    return "\n".join(out)


def build(force: bool = False) -> bool:
    if not os.path.exists(REF_SRC):
        return os.path.exists(PRISTINE)
    os.makedirs(OUT, exist_ok=True)
    src_m = os.path.getmtime(REF_SRC)
    me_m = os.path.getmtime(__file__)
    fresh = all(os.path.exists(p) and os.path.getmtime(p) >= max(src_m, me_m) for p in (PRISTINE, HOOKED))
    if fresh and not force:
        return True
    flags = ["-O2", "-fPIC", "-shared", "-w"]
    subprocess.check_call(["gcc", *flags, "-o", PRISTINE, REF_SRC, "-lz", "-lm"])
    p = subprocess.run(["gcc", *flags, "-x", "c", "-o", HOOKED, "-", "-lz", "-lm"], input=patched_source().encode())
    if p.returncode != 0:
        raise RuntimeError("hooked reference build failed")
    return True


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("reference built" if ok else "reference unavailable (no /root/reference and no prebuilt oracle/_ref)")
