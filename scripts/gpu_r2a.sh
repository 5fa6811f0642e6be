#!/bin/bash
# round 2, call A: full GPU test suite + A/B of the v8 dense_dmma kernel (PDL on/off, flush on/off)
set -u
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest -m gpu" ; timeout 1200 python -m pytest tests -q -m gpu -x > $OUT/r02a_pytest_gpu.log 2>&1 ; echo "exit $?" ; tail -15 $OUT/r02a_pytest_gpu.log
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1 ; echo "exit $?" ; tail -3 $OUT/smoke.log
for v in "" "--no-pdl" "--no-l2-flush" "--no-l2-flush --no-pdl" "--no-l2-flush --dmma-group 2" "--no-l2-flush --dmma-group 16"; do
  echo "== bench $v"
  timeout 600 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-microbench $v 2>> $OUT/bench.err | tee -a $OUT/r02a_bench_ab.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('  value %.4g  ms/step %.4f  e2e %.4g  launches %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches']))"
done
tail -5 $OUT/bench.err
