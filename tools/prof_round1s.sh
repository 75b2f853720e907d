# final pass: GPU tests, cfg2 line with the final generator, cfg2mc with one CTA (default) and two CTAs half an iteration apart
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r1s_pytest_gpu.log 2>&1; tail -2 gpurun_out/r1s_pytest_gpu.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r1s_bench_cfg2.json 2> gpurun_out/r1s_bench_cfg2.err || echo FAIL cfg2
python bench.py --workload cfg2mc --steps 5 --warmup 3 > gpurun_out/r1s_bench_cfg2mc.json 2> gpurun_out/r1s_bench_cfg2mc.err || echo FAIL cfg2mc
SPICEY_JIT_CFG=64,2,112,4,8,0 SPICEY_JIT_ANTIPHASE=14000 python bench.py --workload cfg2mc --steps 5 --warmup 3 > gpurun_out/r1s_bench_cfg2mc_2x64_ap14.json 2>/dev/null || echo FAIL a
SPICEY_JIT_CFG=64,2,112,4,8,0 SPICEY_JIT_ANTIPHASE=0 python bench.py --workload cfg2mc --steps 5 --warmup 3 > gpurun_out/r1s_bench_cfg2mc_2x64_ap0.json 2>/dev/null || echo FAIL b
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r1s_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["value"], d["kernel_ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"])
    except Exception as e: print(f, 'ERR', e)
PY
