#!/bin/bash
# ncu launch list + full capture of the dominant kernel of the default bench.
# Usage: bash scripts/gpu_prof.sh <tag> [kernel-regex] [extra bench args...]
set -u
TAG=${1:-prof}; shift || true
KRE=${1:-half_step}; shift || true
OUT=gpurun_out; mkdir -p $OUT
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/${TAG}_launches.csv \
   python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-l2-flush --no-microbench "$@" > $OUT/${TAG}_ncu_bench.log 2>&1 ; echo "launch list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KRE -s 8 -c 2 -f -o $OUT/${TAG}_full \
   python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-l2-flush --no-microbench "$@" > $OUT/${TAG}_ncu_full.log 2>&1 ; echo "full exit $?"
ls -la $OUT | grep $TAG
