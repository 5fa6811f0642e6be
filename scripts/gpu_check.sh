#!/bin/bash
# One gpurun call: tests, smoke, micro-benchmarks, bench, ncu launch list + full capture.
# Usage (from the repo root on the GPU box):  bash scripts/gpu_check.sh [quick]
set -u
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1 ; echo "exit $?" ; tail -5 $OUT/pytest_gpu.log
echo "== sanitizer" ; timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden_chain" > $OUT/sanitizer.log 2>&1 ; echo "exit $?" ; tail -4 $OUT/sanitizer.log
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1 ; echo "exit $?" ; tail -3 $OUT/smoke.log
echo "== microbench" ; timeout 300 python - > $OUT/microbench.log 2>&1 <<'PY'
from emcee_b200 import _lib
names = {0: "DFMA", 1: "DMMA m8n8k4", 2: "DMMA m16n8k8", 3: "DMMA m16n8k16", 4: "HBM copy"}
for what in (0, 1, 2, 3):
    for wps in (4, 8, 16, 32, 64):
        print("%-14s warps/SM=%2d  %.2f TFLOP/s" % (names[what], wps, _lib.microbench(what, wps)), flush=True)
print("%-14s %.1f GB/s" % (names[4], _lib.microbench(4, 16)))
PY
cat $OUT/microbench.log
echo "== bench" ; timeout 900 python bench.py --steps 100 --warmup 10 > $OUT/bench.json 2> $OUT/bench.err ; echo "exit $?" ; cat $OUT/bench.json ; tail -3 $OUT/bench.err
timeout 600 python bench.py --steps 100 --warmup 10 --no-l2-flush --no-cpu-baseline > $OUT/bench_noflush.json 2>> $OUT/bench.err ; cat $OUT/bench_noflush.json
if [ "${1:-}" != "quick" ]; then
echo "== other workloads"
for wl in "ring 262144 32" "rosenbrock 16384 256" "gauss_dense 4096 128" "gauss_iso 65536 128"; do
  set -- $wl
  timeout 600 python bench.py --workload $1 --nwalkers $2 --ndim $3 --steps 100 --warmup 10 --no-cpu-baseline >> $OUT/bench_other.json 2>> $OUT/bench.err
done
cat $OUT/bench_other.json
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches.csv \
   python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-l2-flush > $OUT/ncu_bench.log 2>&1 ; echo "exit $?"
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:half_step -s 6 -c 2 -f -o $OUT/prof_halfstep \
   python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-l2-flush > $OUT/ncu_full.log 2>&1 ; echo "exit $?"
ls -la $OUT
fi
