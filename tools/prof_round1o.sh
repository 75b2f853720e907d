# cfg2: dependent-chain variants of the compiled sparse kernel (SPICEY_JIT_CHAIN bits), kernel ms per sweep
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for c in ${CHAINS:-0 1 2 3}; do
  SPICEY_JIT_CHAIN=$c python tools/jit_sweep.py "192,1,75,4" > gpurun_out/chain_$c.log 2>&1
  echo "chain=$c: $(tail -2 gpurun_out/chain_$c.log | tr '\n' ' ')"
done
