for dbg in 0 1 2; do
echo "DBG=$dbg"; FGT_COMMIT_DBG=$dbg python bench.py --steps 4 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['device_ms_per_step']['ms_accum'])"
done
