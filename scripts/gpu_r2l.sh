#!/bin/bash
# round 2, call L (1 GPU): tma_rows own-rows-in-registers variant: tests + A/B + ncu + integration test
set -u
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest" ; timeout 1200 python -m pytest tests -q -m gpu > $OUT/r02l_pytest.log 2>&1 ; echo "exit $?" ; tail -4 $OUT/r02l_pytest.log
B="python bench.py --no-cpu-baseline --no-microbench --no-configs --no-l2-flush"
show='import sys, json
for l in sys.stdin:
    d = json.loads(l); print("  %-30s value %.4g  ms/step %.4f  kernel %s frac %.3f" % (d["config"]["workload"][:30], d["value"], d["ms_per_step"], d["kernel"], d["roofline"]["frac"]))'
echo "== bench lines (L2 warm): own_reg on / off"
for opt in "" "--no-own-reg"; do
for wl in "ring 262144 32" "gauss_iso 65536 64" "ring 32768 32" "gauss_iso 262144 32"; do
  set -- $wl
  timeout 300 $B --workload $1 --nwalkers $2 --ndim $3 --steps 100 --warmup 10 $opt 2>>$OUT/bench.err | tee -a $OUT/r02l_bench_hbm.jsonl | python -c "$show"
done
done
echo "== flushed"
BF="python bench.py --no-cpu-baseline --no-microbench --no-configs"
timeout 300 $BF --workload ring --nwalkers 262144 --ndim 32 --steps 100 --warmup 10 2>>$OUT/bench.err | tee -a $OUT/r02l_bench_hbm.jsonl | python -c "$show"
echo "== ncu ring"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:half_step -s 10 -c 2 -f -o $OUT/r02l_ring_tma \
   $B --workload ring --nwalkers 262144 --ndim 32 --steps 8 --warmup 3 > $OUT/r02l_ncu1.log 2>&1 ; echo "exit $?"
tail -3 $OUT/bench.err
