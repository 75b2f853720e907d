cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q > gpurun_out/r1m_pytest_gpu.log 2>&1; tail -3 gpurun_out/r1m_pytest_gpu.log
for w in cfg3 cfg5; do
  python bench.py --workload $w --steps 5 --warmup 3 --device-waves > gpurun_out/r1m_bench_${w}_devwaves.json 2> gpurun_out/r1m_bench_${w}_devwaves.err || echo FAIL $w
  python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/r1m_bench_${w}.json 2> gpurun_out/r1m_bench_${w}.err || echo FAIL $w
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r1m_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["value"], d["kernel_ms_per_step"], d["roofline"]["frac"], d["config"].get("sources"))
    except Exception as e: print(f, 'ERR', e)
PY
