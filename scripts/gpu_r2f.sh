#!/bin/bash
# round 2, call F (2 GPUs): full GPU test-suite (new moves, multi-GPU parity), hybrid local-first A/B
set -u
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest (all gpu tests incl. multi-GPU)" ; timeout 1700 python -m pytest tests -q -m gpu > $OUT/r02f_pytest.log 2>&1 ; echo "exit $?" ; tail -8 $OUT/r02f_pytest.log ; grep -c PASS $OUT/multigpu_check_world2.log; tail -3 $OUT/multigpu_check_world2.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== A/B two GPUs"
port=29900
for sc in weak strong; do
 for fl in 0 1; do
  port=$((port+1)); timeout 300 $TR --master-port $port scripts/ab_step.py --scaling $sc --flush $fl --local-first 0 --tag v11-natural 2>>$OUT/ab.err | grep '^{' | tee -a $OUT/r02f_ab.jsonl
  port=$((port+1)); timeout 300 $TR --master-port $port scripts/ab_step.py --scaling $sc --flush $fl --local-first 1 --tag v11-localfirst 2>>$OUT/ab.err | grep '^{' | tee -a $OUT/r02f_ab.jsonl
 done
done
port=$((port+1)); timeout 300 $TR --master-port $port scripts/timeline_mg.py weak 1 2>>$OUT/ab.err | grep -v "^NCCL" > $OUT/r02f_timeline_mg_weak_lf1.txt ; tail -6 $OUT/r02f_timeline_mg_weak_lf1.txt
tail -3 $OUT/ab.err
