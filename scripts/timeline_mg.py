#!/usr/bin/env python
"""dense_dmma consumer timeline on rank 0 of a multi-GPU (P2P) run: torchrun ... scripts/timeline_mg.py"""
import sys
sys.path.insert(0, ".")
import bench
import emcee_b200
from emcee_b200 import dist, models

rdv = dist.Rendezvous("gloo")
N = 65536 * rdv.world
w = bench.make_workload("gauss_dense", N, 128)
s = emcee_b200.EnsembleSampler(N, 128, models.GaussianDense(w["icov"]), seed=1, device=rdv.local_rank)
eng = s._engine
dist.attach(eng, rdv, "p2p")
eng.set_state(w["p0"])
sched = s._schedule()
eng.step(sched, 20, want_accepted=False)
eng.set_option("dmma_timeline", 1)
rdv.barrier()
eng.step(sched, 3, want_accepted=False)
tl = eng.debug_timeline()
if rdv.rank == 0:
    valid = tl[..., 5] > 0
    ntile = valid.sum(-1)
    print("world", rdv.world, "tiles per SM min/max", ntile.sum(1).min(), ntile.sum(1).max())
    for sm in (0, 100):
        for c in (0, 5):
            rows = []
            for k in range(int(ntile[sm, c])):
                e = tl[sm, c, k]
                rows.append("t%d[%6d w%6d q%5d m%6d e%6d]" % (k, e[1], e[2] - e[1], e[3] - e[2], e[4] - e[3], e[5] - e[4]))
            print("SM%d c%d %s" % (sm, c, " ".join(rows)))
    end = tl[..., 5].max(axis=(1, 2))
    print("kernel end per SM: mean %.0f max %.0f" % (end.mean(), end.max()))
    for k in range(4):
        v = valid[:, :, k]
        if v.any():
            print("tile %d: wait %.0f mma %.0f" % (k, (tl[..., k, 2] - tl[..., k, 1])[v].mean(), (tl[..., k, 4] - tl[..., k, 3])[v].mean()))
rdv.barrier()
rdv.close()
