#!/bin/bash
# round 2, call J (1 GPU): tests; HBM bench lines; sanitizer re-run; full driver-form bench line + reference arm; ncu of the headline
set -u
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest" ; timeout 1200 python -m pytest tests -q -m gpu > $OUT/r02j_pytest.log 2>&1 ; echo "exit $?" ; tail -4 $OUT/r02j_pytest.log
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1 ; echo "exit $?" ; tail -3 $OUT/smoke.log
B="python bench.py --no-cpu-baseline --no-microbench --no-configs --no-l2-flush"
show='import sys, json
for l in sys.stdin:
    d = json.loads(l); print("  %-30s value %.4g  ms/step %.4f  kernel %s frac %.3f" % (d["config"]["workload"][:30], d["value"], d["ms_per_step"], d["kernel"], d["roofline"]["frac"]))'
echo "== bench lines (L2 warm)"
for wl in "ring 262144 32" "rosenbrock 16384 256" "gauss_iso 65536 128" "gauss_iso 65536 64"; do
  set -- $wl
  timeout 300 $B --workload $1 --nwalkers $2 --ndim $3 --steps 100 --warmup 10 2>>$OUT/bench.err | tee -a $OUT/r02j_bench_hbm.jsonl | python -c "$show"
done
echo "== flushed"
BF="python bench.py --no-cpu-baseline --no-microbench --no-configs"
for wl in "ring 262144 32" "rosenbrock 16384 256"; do
  set -- $wl
  timeout 300 $BF --workload $1 --nwalkers $2 --ndim $3 --steps 100 --warmup 10 2>>$OUT/bench.err | tee -a $OUT/r02j_bench_hbm.jsonl | python -c "$show"
done
echo "== sanitizer: memcheck on the new moves + analysis"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_moves_extra.py tests/test_gpu_analysis.py -q -m gpu -x -k "not at_scale and not fixture" > $OUT/r02j_sanitizer_memcheck.log 2>&1 ; echo "exit $?" ; tail -3 $OUT/r02j_sanitizer_memcheck.log
echo "== full bench line (driver form)" ; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/r02j_bench_full.json 2>> $OUT/bench.err ; echo "exit $?" ; python -c "
import json
d=json.loads(open('$OUT/r02j_bench_full.json').read())
print('value %.4g ms %.4f e2e %.4g frac %.3f frac_contract %.3f cpu %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['frac_contract'], d['cpu_baseline']['value']))
for k,v in d['configs'].items(): print('  ', k, v.get('value'), v.get('kernel'), v.get('roofline',{}).get('frac'), v.get('error'))"
echo "== reference arm" ; timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/r02j_bench_reference.json 2>> $OUT/bench.err ; echo "exit $?" ; head -c 600 $OUT/r02j_bench_reference.json ; echo
echo "== default bench (200 steps) + noflush"
timeout 600 python bench.py --no-cpu-baseline --no-configs > $OUT/r02j_bench_200.json 2>>$OUT/bench.err ; python scripts/show_bench.py $OUT/r02j_bench_200.json
timeout 600 python bench.py --no-cpu-baseline --no-configs --no-l2-flush > $OUT/r02j_bench_200_noflush.json 2>>$OUT/bench.err ; python scripts/show_bench.py $OUT/r02j_bench_200_noflush.json
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/r02j_launches.csv \
   python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-l2-flush --no-microbench --no-configs > $OUT/r02j_ncu_bench.log 2>&1 ; echo "exit $?"
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:half_step_dense -s 10 -c 2 -f -o $OUT/r02j_dmma_full \
   python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-l2-flush --no-microbench --no-configs > $OUT/r02j_ncu_full.log 2>&1 ; echo "exit $?"
tail -3 $OUT/bench.err
