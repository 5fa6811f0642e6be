#!/bin/bash
# round 2, call I (1 GPU): tma_rows register path; tests; bench; ncu; compute-sanitizer (memcheck, racecheck, synccheck)
set -u
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest" ; timeout 1200 python -m pytest tests -q -m gpu > $OUT/r02i_pytest.log 2>&1 ; echo "exit $?" ; tail -6 $OUT/r02i_pytest.log
B="python bench.py --no-cpu-baseline --no-microbench --no-configs --no-l2-flush"
show='import sys, json
for l in sys.stdin:
    d = json.loads(l); print("  %-30s value %.4g  ms/step %.4f  kernel %s frac %.3f" % (d["config"]["workload"][:30], d["value"], d["ms_per_step"], d["kernel"], d["roofline"]["frac"]))'
echo "== bench lines (L2 warm)"
for wl in "ring 262144 32" "rosenbrock 16384 256" "gauss_iso 65536 128" "ring 32768 32" "gauss_iso 65536 64"; do
  set -- $wl
  timeout 300 $B --workload $1 --nwalkers $2 --ndim $3 --steps 100 --warmup 10 2>>$OUT/bench.err | tee -a $OUT/r02i_bench_hbm.jsonl | python -c "$show"
done
echo "== flushed"
BF="python bench.py --no-cpu-baseline --no-microbench --no-configs"
for wl in "ring 262144 32" "rosenbrock 16384 256"; do
  set -- $wl
  timeout 300 $BF --workload $1 --nwalkers $2 --ndim $3 --steps 100 --warmup 10 2>>$OUT/bench.err | tee -a $OUT/r02i_bench_hbm.jsonl | python -c "$show"
done
echo "== ncu ring tma_rows"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:half_step -s 10 -c 2 -f -o $OUT/r02i_ring_tma \
   $B --workload ring --nwalkers 262144 --ndim 32 --steps 8 --warmup 3 > $OUT/r02i_ncu1.log 2>&1 ; echo "exit $?"
echo "== ncu rosenbrock tma_rows (DE + snooker)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:half_step -s 10 -c 12 -f -o $OUT/r02i_rosen_tma \
   $B --workload rosenbrock --nwalkers 16384 --ndim 256 --steps 8 --warmup 3 > $OUT/r02i_ncu2.log 2>&1 ; echo "exit $?"
echo "== sanitizer: memcheck on the new moves + analysis"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_moves_extra.py tests/test_gpu_analysis.py -q -m gpu -x -k "not at_scale and not fixture" > $OUT/r02i_sanitizer_memcheck.log 2>&1 ; echo "exit $?" ; tail -3 $OUT/r02i_sanitizer_memcheck.log
echo "== sanitizer: racecheck (dense_dmma, tma_rows, generic goldens)"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "golden_run_mcmc_bulk and (stretch_dense_64x8 or stretch_ring_80x6 or mix_de_snooker or stretch_iso_32x5)" > $OUT/r02i_sanitizer_racecheck.log 2>&1 ; echo "exit $?" ; tail -3 $OUT/r02i_sanitizer_racecheck.log
echo "== sanitizer: synccheck"
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "golden_run_mcmc_bulk and (stretch_dense_64x8 or stretch_ring_80x6 or mix_de_snooker or walk_all_dense)" > $OUT/r02i_sanitizer_synccheck.log 2>&1 ; echo "exit $?" ; tail -3 $OUT/r02i_sanitizer_synccheck.log
tail -3 $OUT/bench.err
