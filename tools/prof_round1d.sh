set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q > gpurun_out/r1k_pytest_gpu.log 2>&1; tail -2 gpurun_out/r1k_pytest_gpu.log
for w in cfg1 cfg2 cfg2mc cfg3 cfg5; do
  python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/r1k_bench_$w.json 2> gpurun_out/r1k_bench_$w.err || echo FAIL $w
done
python bench.py --workload cfg4 --points 400000 --steps 3 --warmup 3 > gpurun_out/r1k_bench_cfg4.json 2> gpurun_out/r1k_bench_cfg4.err || echo FAIL cfg4
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1k_ref_cfg2.json 2>/dev/null
python bench.py --steps 2 --warmup 3 > gpurun_out/plain_cfg2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1k_launches_cfg2.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_l2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spicey_sparse_jit --launch-skip 3 -c 1 -o gpurun_out/prof_cfg2_jit_k -f python bench.py --workload cfg2 --steps 2 --warmup 3 > gpurun_out/ncu_f2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ac_warp_kernel --launch-skip 1 -c 1 -o gpurun_out/prof_cfg4_warp_k -f python bench.py --workload cfg4 --points 60000 --steps 1 --warmup 1 > gpurun_out/ncu_f4.log 2>&1
ls -la gpurun_out | grep r1k
