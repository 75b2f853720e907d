# 8-GPU lines (torchrun, one rank per GPU, NUMA-bound ranks): cfg2 (headline), cfg4 slice, cfg5
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r1n_bench_cfg2_n8.json 2> gpurun_out/r1n_bench_cfg2_n8.err
$TR --master-port 29522 bench.py --gpus 8 --steps 3 --warmup 3 --workload cfg4 --points 400000 > gpurun_out/r1n_bench_cfg4_n8.json 2> gpurun_out/r1n_bench_cfg4_n8.err
$TR --master-port 29523 bench.py --gpus 8 --steps 3 --warmup 3 --workload cfg5 > gpurun_out/r1n_bench_cfg5_n8.json 2> gpurun_out/r1n_bench_cfg5_n8.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r1n_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["value"], d["kernel_ms_per_step"], d["e2e"]["value"], d["e2e"]["pcie_gbs"], d["e2e"].get("host_numa"))
    except Exception as e: print(f, 'ERR', e)
PY
