#!/bin/bash
# round 2, call E (2 GPUs): locality-sorted tiles + deferred peer barrier; full GPU test-suite incl. multi-GPU parity
set -u
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest (all gpu tests incl. multi-GPU)" ; timeout 1700 python -m pytest tests -q -m gpu > $OUT/r02e_pytest.log 2>&1 ; echo "exit $?" ; tail -12 $OUT/r02e_pytest.log ; grep -c PASS $OUT/multigpu_check_world2.log; tail -4 $OUT/multigpu_check_world2.log
R1=$PWD/emcee_b200/libemcee_b200_r1.so
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== A/B single GPU"
for fl in 0 1; do
  EMCEE_B200_LIB=$R1 timeout 300 python scripts/ab_step.py --flush $fl --tag r1 2>>$OUT/ab.err | tee -a $OUT/r02e_ab.jsonl
  timeout 300 python scripts/ab_step.py --flush $fl --tag v10 2>>$OUT/ab.err | tee -a $OUT/r02e_ab.jsonl
  timeout 300 python scripts/ab_step.py --flush $fl --pdl 0 --tag v10-nopdl 2>>$OUT/ab.err | tee -a $OUT/r02e_ab.jsonl
done
echo "== A/B two GPUs"
port=29800
for sc in weak strong; do
 for fl in 0 1; do
  port=$((port+1)); timeout 300 $TR --master-port $port scripts/ab_step.py --scaling $sc --flush $fl --local-first 0 --tag v10-natural 2>>$OUT/ab.err | grep '^{' | tee -a $OUT/r02e_ab.jsonl
  port=$((port+1)); timeout 300 $TR --master-port $port scripts/ab_step.py --scaling $sc --flush $fl --local-first 1 --tag v10-localfirst 2>>$OUT/ab.err | grep '^{' | tee -a $OUT/r02e_ab.jsonl
 done
 port=$((port+1)); timeout 300 $TR --master-port $port scripts/ab_step.py --scaling $sc --local-first 1 --group 2 --tag v10-lf-group2 2>>$OUT/ab.err | grep '^{' | tee -a $OUT/r02e_ab.jsonl
done
echo "== timelines 2 GPUs"
for lf in 0 1; do port=$((port+1)); timeout 300 $TR --master-port $port scripts/timeline_mg.py weak $lf 2>>$OUT/ab.err | grep -v "^NCCL" > $OUT/r02e_timeline_mg_weak_lf$lf.txt ; tail -6 $OUT/r02e_timeline_mg_weak_lf$lf.txt; done
port=$((port+1)); timeout 300 $TR --master-port $port scripts/timeline_mg.py strong 1 2>>$OUT/ab.err | grep -v "^NCCL" > $OUT/r02e_timeline_mg_strong_lf1.txt ; tail -6 $OUT/r02e_timeline_mg_strong_lf1.txt
echo "== bench --gpus 2" ; port=$((port+1)); timeout 900 $TR --master-port $port bench.py --gpus 2 --steps 100 --warmup 10 --no-configs 2> $OUT/r02e_bench_g2.err | grep '^{' > $OUT/r02e_bench_g2.json ; python -c "
import json
d=json.loads(open('$OUT/r02e_bench_g2.json').read())
print('strong value %.4g ms %.4f e2e %.4g parity %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['parity_checked']))
w=d['weak']; print('weak value %.4g ms %.4f e2e %.4g parity %s' % (w['value'], w['ms_per_step'], w['e2e']['value'], w['parity_checked']))"
tail -5 $OUT/ab.err
