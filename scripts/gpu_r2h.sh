#!/bin/bash
# round 2, call H (1 GPU): tma_rows with per-lane batch preparation (+ long rows); tests; bench A/B; ncu
set -u
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest" ; timeout 1200 python -m pytest tests -q -m gpu -x > $OUT/r02h_pytest.log 2>&1 ; echo "exit $?" ; tail -6 $OUT/r02h_pytest.log
B="python bench.py --no-cpu-baseline --no-microbench --no-configs --no-l2-flush"
show='import sys, json
for l in sys.stdin:
    d = json.loads(l); print("  %-30s value %.4g  ms/step %.4f  kernel %s frac %.3f" % (d["config"]["workload"][:30], d["value"], d["ms_per_step"], d["kernel"], d["roofline"]["frac"]))'
echo "== bench lines (L2 warm)"
for opt in 2 1; do
for wl in "ring 262144 32" "rosenbrock 16384 256" "gauss_iso 65536 128" "ring 32768 32"; do
  set -- $wl
  timeout 300 $B --workload $1 --nwalkers $2 --ndim $3 --steps 100 --warmup 10 --tma-rows $opt 2>>$OUT/bench.err | tee -a $OUT/r02h_bench_hbm.jsonl | python -c "$show"
done
done
timeout 300 $B --workload rosenbrock --nwalkers 16384 --ndim 256 --steps 100 --warmup 10 --no-tma-rows 2>>$OUT/bench.err | tee -a $OUT/r02h_bench_hbm.jsonl | python -c "$show"
echo "== flushed"
BF="python bench.py --no-cpu-baseline --no-microbench --no-configs"
for wl in "ring 262144 32" "rosenbrock 16384 256"; do
  set -- $wl
  timeout 300 $BF --workload $1 --nwalkers $2 --ndim $3 --steps 100 --warmup 10 2>>$OUT/bench.err | tee -a $OUT/r02h_bench_hbm.jsonl | python -c "$show"
done
echo "== ncu ring tma_rows"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:half_step -s 10 -c 2 -f -o $OUT/r02h_ring_tma \
   $B --workload ring --nwalkers 262144 --ndim 32 --steps 8 --warmup 3 > $OUT/r02h_ncu1.log 2>&1 ; echo "exit $?"
echo "== ncu rosenbrock tma_rows long rows (DE + snooker)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:half_step -s 10 -c 12 -f -o $OUT/r02h_rosen_tma \
   $B --workload rosenbrock --nwalkers 16384 --ndim 256 --steps 8 --warmup 3 > $OUT/r02h_ncu2.log 2>&1 ; echo "exit $?"
tail -3 $OUT/bench.err
