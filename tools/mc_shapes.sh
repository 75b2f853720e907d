#!/bin/bash
# GPU: bench.py --workload cfg2mc (AC Monte-Carlo through the compiled tier with per-instance stamping) for several
# launch shapes: each argument "block,CTAs,smem slots,barrier period,prefetch:register values" (or "default:").
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
o=gpurun_out/${MC_TAG:-mc_shapes}.txt; : > $o
for v in "$@"; do
  cfg=${v%%:*}; rv=${v##*:}
  echo "# $v" >> $o
  if [ "$cfg" = default ]; then unset SPICEY_JIT_CFG SPICEY_JIT_REGVALUES; else export SPICEY_JIT_CFG=$cfg SPICEY_JIT_REGVALUES=$rv; fi
  timeout 150 python bench.py --workload cfg2mc --steps 5 --warmup 3 2>>$o.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('%.1f M solves/s  %.3f ms  hbm frac %.3f  tier %d  relerr %.2e' % (d['value'] / 1e6, d['ms_per_step'], d['roofline']['frac'], d['details']['tier'], d['details']['checked_against_oracle']['max_rel_err']))" >> $o 2>&1
done
cat $o
