set -x
cd $GRAFT_REPO_ROOT
for w in cfg2 cfg3 cfg5; do
  python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/r1b_bench_$w.json 2> gpurun_out/r1b_bench_$w.err || echo FAIL $w
done
python bench.py --steps 2 --warmup 3 > gpurun_out/plain_cfg2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1b_launches_cfg2.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_l2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spicey_sparse_jit --launch-skip 3 -c 1 -o gpurun_out/prof_cfg2_jit_final -f python bench.py --workload cfg2 --steps 2 --warmup 3 > gpurun_out/ncu_f2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spicey_tran_jit --launch-skip 2 -c 1 -o gpurun_out/prof_cfg3_jit -f python bench.py --workload cfg3 --steps 2 --warmup 3 > gpurun_out/ncu_f3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spicey_tran_jit --launch-skip 2 -c 1 -o gpurun_out/prof_cfg5_jit -f python bench.py --workload cfg5 --steps 2 --warmup 3 > gpurun_out/ncu_f5.log 2>&1
ls -la gpurun_out | tail -12
