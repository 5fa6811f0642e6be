#!/bin/bash
# round 2, call K (8 GPUs): multi-GPU parity at world 8 (quick), driver-form bench at 8 GPUs, strong-scaling A/B
set -u
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L > $OUT/r02k${NG:-8}_gpus.txt 2>&1
echo "== multi-GPU parity, world 8 (quick)"; EB_MG_WORLD=${NG:-8} EB_MG_QUICK=1 timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu > $OUT/r02k${NG:-8}_pytest.log 2>&1 ; echo "exit $?" ; tail -3 $OUT/r02k${NG:-8}_pytest.log ; grep -c PASS $OUT/multigpu_check_world${NG:-8}.log ; tail -2 $OUT/multigpu_check_world${NG:-8}.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-8} --master-addr 127.0.0.1"
echo "== bench --gpus ${NG:-8} (driver form)"
timeout 900 $TR --master-port 29951 bench.py --gpus ${NG:-8} --steps 20 --warmup 5 2> $OUT/r02k${NG:-8}_bench_g8.err | grep '^{' > $OUT/r02k${NG:-8}_bench_g8.json ; echo "exit $?"; python -c "
import json
d=json.loads(open('$OUT/r02k${NG:-8}_bench_g8.json').read())
print('strong value %.4g ms %.4f e2e %.4g parity %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['parity_checked']))
w=d['weak']; print('weak value %.4g ms %.4f e2e %.4g parity %s' % (w['value'], w['ms_per_step'], w['e2e']['value'], w['parity_checked']))
for k,v in (d.get('configs') or {}).items(): print('  ', k, v.get('value'), v.get('kernel'), v.get('error'))"
echo "== reference arm under torchrun"
timeout 600 $TR --master-port 29952 bench.py --impl reference --gpus ${NG:-8} --steps 20 --warmup 5 2>> $OUT/r02k${NG:-8}_bench_g8.err | grep '^{' | head -c 300; echo
echo "== A/B 8 GPUs"
port=29960
for sc in strong weak; do
 for lf in 0 2; do
  port=$((port+1)); timeout 300 $TR --master-port $port scripts/ab_step.py --scaling $sc --local-first $lf --tag v11-lf$lf 2>>$OUT/ab.err | grep '^{' | tee -a $OUT/r02k${NG:-8}_ab.jsonl
 done
done
port=$((port+1)); timeout 300 $TR --master-port $port scripts/timeline_mg.py weak 0 2>>$OUT/ab.err | grep -v "^NCCL" > $OUT/r02k${NG:-8}_timeline_mg_weak.txt ; tail -6 $OUT/r02k${NG:-8}_timeline_mg_weak.txt
tail -3 $OUT/ab.err; tail -3 $OUT/r02k${NG:-8}_bench_g8.err
