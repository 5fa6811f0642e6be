#!/bin/bash
# round 2, call D (2 GPUs): restored v7 pipeline + PDL/abort; multi-GPU parity; new moves; A/B vs r1 library
set -u
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest (all gpu tests incl. multi-GPU)" ; timeout 1500 python -m pytest tests -q -m gpu -x > $OUT/r02d_pytest.log 2>&1 ; echo "exit $?" ; tail -12 $OUT/r02d_pytest.log ; grep -c PASS $OUT/multigpu_check_world2.log; tail -4 $OUT/multigpu_check_world2.log
R1=$PWD/emcee_b200/libemcee_b200_r1.so
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== A/B single GPU"
for fl in 0 1; do
  EMCEE_B200_LIB=$R1 timeout 300 python scripts/ab_step.py --flush $fl --tag r1 2>>$OUT/ab.err | tee -a $OUT/r02d_ab.jsonl
  timeout 300 python scripts/ab_step.py --flush $fl --tag v9 2>>$OUT/ab.err | tee -a $OUT/r02d_ab.jsonl
  timeout 300 python scripts/ab_step.py --flush $fl --pdl 0 --tag v9-nopdl 2>>$OUT/ab.err | tee -a $OUT/r02d_ab.jsonl
done
echo "== A/B two GPUs"
port=29700
for sc in weak strong; do
  port=$((port+1)); EMCEE_B200_LIB=$R1 timeout 300 $TR --master-port $port scripts/ab_step.py --scaling $sc --tag r1 2>>$OUT/ab.err | grep '^{' | tee -a $OUT/r02d_ab.jsonl
  port=$((port+1)); timeout 300 $TR --master-port $port scripts/ab_step.py --scaling $sc --tag v9 2>>$OUT/ab.err | grep '^{' | tee -a $OUT/r02d_ab.jsonl
  port=$((port+1)); timeout 300 $TR --master-port $port scripts/ab_step.py --scaling $sc --pdl 2 --tag v9-pdl2 2>>$OUT/ab.err | grep '^{' | tee -a $OUT/r02d_ab.jsonl
  port=$((port+1)); timeout 300 $TR --master-port $port scripts/ab_step.py --scaling $sc --group 2 --tag v9-group2 2>>$OUT/ab.err | grep '^{' | tee -a $OUT/r02d_ab.jsonl
done
echo "== timeline 1 GPU" ; timeout 300 python scripts/timeline.py > $OUT/r02d_timeline.txt 2>&1 ; tail -7 $OUT/r02d_timeline.txt
echo "== bench --gpus 2" ; port=$((port+1)); timeout 900 $TR --master-port $port bench.py --gpus 2 --steps 100 --warmup 10 --no-configs > $OUT/r02d_bench_g2.json 2> $OUT/r02d_bench_g2.err ; echo "exit $?" ; python scripts/show_bench.py $OUT/r02d_bench_g2.json 2>/dev/null | head -20
tail -5 $OUT/ab.err
