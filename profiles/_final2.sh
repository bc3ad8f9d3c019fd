# final captures of the round-2 build (second half): bench line, launch list, --set full of the accumulation kernels
python bench.py > gpurun_out/r2c_bench_n1.json 2> gpurun_out/bench.err
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 800 --csv --log-file gpurun_out/r2c_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_eval_pairs|k_commit_ordered" -s 2 -c 2 -o gpurun_out/r2c_full -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-dense > gpurun_out/ncu2.log 2>&1
tail -c 600 gpurun_out/r2c_bench_n1.json
