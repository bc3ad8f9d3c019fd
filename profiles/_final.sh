python -m pytest tests -m gpu -x -q > gpurun_out/gputest.log 2>&1; tail -2 gpurun_out/gputest.log
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/bench.err
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 800 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_eval_pairs|k_commit_ordered" -s 2 -c 2 -o gpurun_out/r2_full -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-dense > gpurun_out/ncu2.log 2>&1
N=100 KEY=null_impl SB=2 ncu --set full --clock-control none --import-source on -k regex:"k_null_eval2|k_commit_ordered" -s 6 -c 3 -o gpurun_out/r2_null_full -f python profiles/quick_null.py > gpurun_out/ncu3.log 2>&1
python bench.py --config 3 --scale 0.2 --chrom-frac 0.1 > gpurun_out/cfg3.json 2> gpurun_out/cfg3.err
python bench.py --config 4 --chrom-frac 0.1 > gpurun_out/cfg4.json 2> gpurun_out/cfg4.err
python bench.py --config 5 --scale 0.04 --chrom-frac 0.05 > gpurun_out/cfg5.json 2> gpurun_out/cfg5.err
IMPLS=1,0 python profiles/null_tutorial.py 2>&1 | tail -2
