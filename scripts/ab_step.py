#!/usr/bin/env python
"""Same-box A/B of two builds of the library on the headline workload (old C-ABI subset only, so it runs
against the round-1 library as well):  EMCEE_B200_LIB=<.so> python [-m torch.distributed.run ...] scripts/ab_step.py"""
import argparse
import json
import os
import sys

sys.path.insert(0, ".")
import numpy as np  # noqa: E402

import bench  # noqa: E402
from emcee_b200 import _lib, dist  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scaling", default="strong")
ap.add_argument("--steps", type=int, default=200)
ap.add_argument("--flush", type=int, default=0)
ap.add_argument("--group", type=int, default=0)
ap.add_argument("--pdl", type=int, default=-1)
ap.add_argument("--nwalkers", type=int, default=65536)
ap.add_argument("--local-first", type=int, default=-1)
ap.add_argument("--tag", default="")
a = ap.parse_args()
rdv = dist.Rendezvous()
N = a.nwalkers * (rdv.world if a.scaling == "weak" else 1)
w = bench.make_workload("gauss_dense", N, 128)
eng = _lib.Engine(N, 128, bench.SAMPLER_SEED, device=rdv.local_rank)
eng.set_model("gauss_dense", np.concatenate([np.zeros(128), w["icov"].ravel()]))
dist.attach(eng, rdv, "p2p")
eng.set_option("l2_flush", a.flush)
if a.group > 0:
    eng.set_option("dmma_group", a.group)
if a.pdl >= 0:
    eng.set_option("pdl", a.pdl)
if a.local_first >= 0:
    eng.set_option("dmma_local_first", a.local_first)
sched = [(dict(kind="stretch", nsplits=2, randomize_split=True, live_dangerously=False, p0=2.0, p1=float("nan")), 1.0)]
eng.set_state(w["p0"])
eng.step(sched, 20, want_accepted=False)
best = None
for rep in range(3):
    rdv.barrier()
    eng.step(sched, a.steps, want_accepted=False)
    ms = rdv.max(eng.last_step_timing()[0])
    best = ms if best is None else min(best, ms)
if rdv.rank == 0:
    print(json.dumps({"lib": os.path.basename(os.environ.get("EMCEE_B200_LIB", "current")), "tag": a.tag, "world": rdv.world,
                      "scaling": a.scaling, "N": N, "flush": a.flush, "group": a.group, "pdl": a.pdl,
                      "us_per_step": 1e3 * best / a.steps, "value": N * a.steps / (best * 1e-3)}), flush=True)
eng.close()
rdv.close()
