#!/bin/bash
# quick iteration loop on the GPU box: parity tests, headline bench (flushed / warm L2), optional profile
TAG=${1:-iter}
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x 2>&1 | tail -3
python bench.py --steps 200 --warmup 20 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err
python bench.py --steps 200 --warmup 20 --no-l2-flush --no-cpu-baseline > gpurun_out/bench_noflush.json 2>> gpurun_out/bench.err
if [ "${2:-}" == "prof" ]; then bash scripts/gpu_prof.sh $TAG half_step; fi
