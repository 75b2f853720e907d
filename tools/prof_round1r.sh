# final pass of the round: GPU tests, bench lines (cfg2 default shape = 2 x 96 threads per SM half an iteration apart), launch list
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r1r_pytest_gpu.log 2>&1; tail -2 gpurun_out/r1r_pytest_gpu.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r1r_bench_cfg2.json 2> gpurun_out/r1r_bench_cfg2.err || echo FAIL cfg2
python bench.py --workload cfg2mc --steps 5 --warmup 3 > gpurun_out/r1r_bench_cfg2mc.json 2> gpurun_out/r1r_bench_cfg2mc.err || echo FAIL cfg2mc
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r1r_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["value"], d["kernel_ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"])
    except Exception as e: print(f, 'ERR', e)
PY
