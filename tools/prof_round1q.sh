# cfg2: two / three CTAs per SM started out of phase (one eliminates while the other stores)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for ap in ${APS:-0 6000 9000 12000}; do
  echo "antiphase=$ap: $(SPICEY_JIT_ANTIPHASE=$ap python tools/jit_sweep.py ${CFGS:-96,2,75,4,8,0} 2>&1 | grep cfg= | tr '\n' ' ')"
done
