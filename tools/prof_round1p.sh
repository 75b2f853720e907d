# cfg2: does the compiled kernel's time per SM depend on how many SMs run it (HBM contention in the store phase)?
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for g in 148 111 74 37; do
  echo "grid=$g: $(SPICEY_JIT_GRID=$g python tools/jit_sweep.py '192,1,75,4' 2>&1 | tail -1)"
done
