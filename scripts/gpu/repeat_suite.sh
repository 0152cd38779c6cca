set -x
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
for i in 1 2 3; do
  timeout 1200 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r2c30_pytest_$i.log 2>&1; echo "run $i rc=$?"; tail -1 gpurun_out/r2c30_pytest_$i.log
done
for i in 1 2; do timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2c30_bench_$i.json 2>/dev/null; echo "bench $i rc=$?"; done
python - <<'PY'
import json
for i in (1,2):
    d=json.loads(open(f'gpurun_out/r2c30_bench_{i}.json').read().strip().splitlines()[-1])
    print(i, d['value'], d['roofline']['kernel_ms'], round(d['roofline']['frac'],4), d['parity']['pass'], d['clocks']['sm_mhz'])
PY
