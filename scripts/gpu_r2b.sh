#!/bin/bash
# round 2, call B: v8 kernel with own-row loads hoisted; device analysis tests; full bench line; timeline; ncu
set -u
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest analysis + parity" ; timeout 900 python -m pytest tests/test_gpu_analysis.py tests/test_gpu_parity.py tests/test_gpu_scale.py -q -m gpu -x > $OUT/r02b_pytest.log 2>&1 ; echo "exit $?" ; tail -12 $OUT/r02b_pytest.log
for v in "" "--no-l2-flush" "--no-l2-flush --no-pdl"; do
  echo "== bench $v"
  timeout 600 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-microbench --no-configs $v 2>> $OUT/bench.err | tee -a $OUT/r02b_bench_ab.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('  value %.4g  ms/step %.4f  e2e %.4g  launches %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches']))"
done
echo "== timeline" ; timeout 300 python scripts/timeline.py > $OUT/r02b_timeline.txt 2>&1 ; tail -8 $OUT/r02b_timeline.txt
echo "== full bench line (driver form)" ; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/r02b_bench_full.json 2>> $OUT/bench.err ; echo "exit $?" ; head -c 3000 $OUT/r02b_bench_full.json ; echo
echo "== reference arm" ; timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/r02b_bench_reference.json 2>> $OUT/bench.err ; echo "exit $?" ; head -c 1500 $OUT/r02b_bench_reference.json ; echo
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/r02b_launches.csv \
   python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-l2-flush --no-microbench --no-configs > $OUT/r02b_ncu_bench.log 2>&1 ; echo "exit $?"
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:half_step_dense -s 10 -c 2 -f -o $OUT/r02b_dmma_full \
   python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-l2-flush --no-microbench --no-configs > $OUT/r02b_ncu_full.log 2>&1 ; echo "exit $?"
tail -5 $OUT/bench.err
ls -la $OUT | tail -12
