#!/bin/bash
# round 2, call C (2 GPUs): multi-GPU parity under pytest, same-box A/B r1 library vs current (1 and 2 GPUs), timeline
set -u
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L > $OUT/r02c_gpus.txt 2>&1
echo "== multi-GPU pytest" ; EB_MG_QUICK=${EB_MG_QUICK:-0} timeout 1200 python -m pytest tests/test_gpu_multi.py tests/test_gpu_analysis.py -q -m gpu -x > $OUT/r02c_pytest.log 2>&1 ; echo "exit $?" ; tail -8 $OUT/r02c_pytest.log ; tail -30 $OUT/multigpu_check_world2.log
R1=$PWD/emcee_b200/libemcee_b200_r1.so
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== A/B single GPU"
for fl in 0 1; do
  EMCEE_B200_LIB=$R1 timeout 300 python scripts/ab_step.py --flush $fl --tag r1 2>>$OUT/ab.err | tee -a $OUT/r02c_ab.jsonl
  timeout 300 python scripts/ab_step.py --flush $fl --tag v8 2>>$OUT/ab.err | tee -a $OUT/r02c_ab.jsonl
  timeout 300 python scripts/ab_step.py --flush $fl --pdl 0 --tag v8-nopdl 2>>$OUT/ab.err | tee -a $OUT/r02c_ab.jsonl
done
echo "== A/B two GPUs"
port=29600
for sc in weak strong; do
  port=$((port+1)); EMCEE_B200_LIB=$R1 timeout 300 $TR --master-port $port scripts/ab_step.py --scaling $sc --tag r1 2>>$OUT/ab.err | grep '^{' | tee -a $OUT/r02c_ab.jsonl
  port=$((port+1)); timeout 300 $TR --master-port $port scripts/ab_step.py --scaling $sc --tag v8 2>>$OUT/ab.err | grep '^{' | tee -a $OUT/r02c_ab.jsonl
  port=$((port+1)); timeout 300 $TR --master-port $port scripts/ab_step.py --scaling $sc --pdl 0 --tag v8-nopdl 2>>$OUT/ab.err | grep '^{' | tee -a $OUT/r02c_ab.jsonl
  port=$((port+1)); timeout 300 $TR --master-port $port scripts/ab_step.py --scaling $sc --group 2 --tag v8-group2 2>>$OUT/ab.err | grep '^{' | tee -a $OUT/r02c_ab.jsonl
  port=$((port+1)); timeout 300 $TR --master-port $port scripts/ab_step.py --scaling $sc --group 16 --tag v8-group16 2>>$OUT/ab.err | grep '^{' | tee -a $OUT/r02c_ab.jsonl
done
echo "== timeline 2 GPUs" ; port=$((port+1)); timeout 300 $TR --master-port $port scripts/timeline_mg.py > $OUT/r02c_timeline_mg.txt 2>&1 ; tail -12 $OUT/r02c_timeline_mg.txt
echo "== bench --gpus 2" ; port=$((port+1)); timeout 900 $TR --master-port $port bench.py --gpus 2 --steps 100 --warmup 10 > $OUT/r02c_bench_g2.json 2> $OUT/r02c_bench_g2.err ; echo "exit $?" ; head -c 2500 $OUT/r02c_bench_g2.json ; echo ; tail -5 $OUT/r02c_bench_g2.err
tail -5 $OUT/ab.err
