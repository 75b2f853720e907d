#!/bin/bash
# One profiling recipe for every kernel (replaces the per-experiment scripts of round 1).  Run under gpurun:
#   tools/prof.sh <tag> <kernel-regex> <skip> <count> -- <command...>
# 1. runs <command> plainly (must exit 0), 2. captures <count> launches matching <kernel-regex> after <skip> with
# `ncu --set full --import-source on`, 3. writes gpurun_out/<tag>.ncu-rep, a raw-metric summary <tag>_raw.txt and the
# hottest source lines <tag>_src.txt (tools/ncu_raw.py, tools/ncu_src_top.py).  Copy what should be judged to profiles/.
set -u
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
tag=$1; regex=$2; skip=$3; count=$4; shift 4
[ "$1" = "--" ] && shift
"$@" > gpurun_out/${tag}_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/${tag}_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:${regex} -s ${skip} -c ${count} -f -o gpurun_out/${tag} "$@" > gpurun_out/${tag}_ncu.log 2>&1 || { echo "ncu failed"; tail -20 gpurun_out/${tag}_ncu.log; exit 1; }
python tools/ncu_raw.py gpurun_out/${tag}.ncu-rep > gpurun_out/${tag}_raw.txt 2>&1
ncu -i gpurun_out/${tag}.ncu-rep --page source --csv > gpurun_out/${tag}_source.csv 2>/dev/null
python tools/ncu_src_top.py gpurun_out/${tag}_source.csv 60 > gpurun_out/${tag}_src.txt 2>&1 || true
tail -50 gpurun_out/${tag}_raw.txt
