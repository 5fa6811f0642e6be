#!/bin/bash
# final verification of a round on one B200: full GPU test-suite, smoke, headline bench, ncu evidence
set -u
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 600 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "exit $?"; tail -2 $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "exit $?"; tail -2 $OUT/smoke.log
echo "== bench"; timeout 600 python bench.py > $OUT/final_bench.json 2> $OUT/final_bench.err; echo "exit $?"; tail -2 $OUT/final_bench.err
timeout 600 python bench.py --no-l2-flush --no-cpu-baseline > $OUT/final_bench_noflush.json 2>> $OUT/final_bench.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > $OUT/final_bench_reference.json 2>> $OUT/final_bench.err
echo "== ncu"; bash scripts/gpu_prof.sh r01_final half_step
