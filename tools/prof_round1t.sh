# final coherent set of single-GPU bench lines of round 1 (r1s_*)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for w in cfg1 cfg3 cfg5; do
  python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/r1s_bench_$w.json 2> gpurun_out/r1s_bench_$w.err || echo FAIL $w
done
python bench.py --workload cfg4 --steps 3 --warmup 3 > gpurun_out/r1s_bench_cfg4.json 2> gpurun_out/r1s_bench_cfg4.err || echo FAIL cfg4
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r1s_bench_cfg[1345].json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["value"], d["kernel_ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["config"]["tier"])
    except Exception as e: print(f, 'ERR', e)
PY
