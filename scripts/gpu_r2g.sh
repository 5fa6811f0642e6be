#!/bin/bash
# round 2, call G (1 GPU): ncu of the HBM-bound kernels (ring tma_rows / generic, rosenbrock DE + snooker generic), bench lines
set -u
OUT=gpurun_out
mkdir -p $OUT
B="python bench.py --no-cpu-baseline --no-microbench --no-configs --no-l2-flush"
echo "== bench lines (L2 warm)"
for wl in "ring 262144 32" "rosenbrock 16384 256" "gauss_iso 65536 128"; do
  set -- $wl
  timeout 300 $B --workload $1 --nwalkers $2 --ndim $3 --steps 100 --warmup 10 2>>$OUT/bench.err | tee -a $OUT/r02g_bench_hbm.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('  %s value %.4g  ms/step %.4f  kernel %s frac %.3f' % (d['config']['workload'][:28], d['value'], d['ms_per_step'], d['kernel'], d['roofline']['frac']))"
done
timeout 300 $B --workload ring --nwalkers 262144 --ndim 32 --steps 100 --warmup 10 --no-tma-rows 2>>$OUT/bench.err | tee -a $OUT/r02g_bench_hbm.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('  %s value %.4g  ms/step %.4f  kernel %s frac %.3f' % (d['config']['workload'][:28], d['value'], d['ms_per_step'], d['kernel'], d['roofline']['frac']))"
echo "== ncu ring tma_rows"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:half_step -s 10 -c 2 -f -o $OUT/r02g_ring_tma \
   $B --workload ring --nwalkers 262144 --ndim 32 --steps 8 --warmup 3 > $OUT/r02g_ncu1.log 2>&1 ; echo "exit $?"
echo "== ncu rosenbrock generic (DE + snooker)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:half_step -s 10 -c 12 -f -o $OUT/r02g_rosen_generic \
   $B --workload rosenbrock --nwalkers 16384 --ndim 256 --steps 8 --warmup 3 > $OUT/r02g_ncu2.log 2>&1 ; echo "exit $?"
ls -la $OUT/*.ncu-rep | tail -3
