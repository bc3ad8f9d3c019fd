for w in 0 1; do N=100 K=30 L=100000 python - <<EOF
import os
os.environ["KEY"]="null_impl"; os.environ["SB"]="2"
src=open("profiles/quick_null.py").read().replace("lib = PackedCompanion(fbuild.build())","lib = PackedCompanion(fbuild.build()); lib.set_option(\"null_window\",$w)")
print("window $w", end=" "); exec(src)
EOF
done
IMPLS=0 VERBOSE=1 python profiles/null_tutorial.py 2>&1 | tail -14
python bench.py --config 4 --chrom-frac 0.1 > gpurun_out/cfg4.json 2>gpurun_out/cfg4.err; tail -c 300 gpurun_out/cfg4.json
