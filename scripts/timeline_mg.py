#!/usr/bin/env python
"""dense_dmma consumer + producer timeline on rank 0 of a multi-GPU (P2P) run:
   python -m torch.distributed.run ... scripts/timeline_mg.py [weak|strong] [local_first 0|1]"""
import sys
sys.path.insert(0, ".")
import numpy as np
import bench
import emcee_b200
from emcee_b200 import dist, models

scaling = sys.argv[1] if len(sys.argv) > 1 else "weak"
local_first = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rdv = dist.Rendezvous()
N = 65536 * (rdv.world if scaling == "weak" else 1)
w = bench.make_workload("gauss_dense", N, 128)
s = emcee_b200.EnsembleSampler(N, 128, models.GaussianDense(w["icov"]), seed=1, device=rdv.local_rank)
s.attach(rdv, "p2p")
eng = s._engine
eng.set_option("dmma_local_first", local_first)
eng.set_state(w["p0"])
sched = s._schedule()
eng.step(sched, 20, want_accepted=False)
eng.set_option("dmma_timeline", 1)
rdv.barrier()
eng.step(sched, 3, want_accepted=False)
tl = eng.debug_timeline()  # [SM, pair, tile, event]
if rdv.rank == 0:
    valid = tl[..., 5] > 0
    ntile = valid.sum(-1)
    print("world", rdv.world, scaling, "local_first", local_first, "tiles per SM min/max", ntile.sum(1).min(), ntile.sum(1).max())
    for sm in (0, 100):
        for c in (0, 5):
            rows = []
            for k in range(int(ntile[sm, c])):
                e = tl[sm, c, k]
                rows.append("t%d[%6d w%6d q%5d m%6d e%6d | req%6d land%6d pub%6d]" % (
                    k, e[1], e[2] - e[1], e[3] - e[2], e[4] - e[3], e[5] - e[4], e[6], e[7], e[8]))
            print("SM%d c%d %s" % (sm, c, " ".join(rows)))
    end = tl[..., 5].max(axis=(1, 2))
    print("kernel end per SM: mean %.0f max %.0f" % (end.mean(), end.max()))
    for k in range(4):
        v = valid[:, :, k]
        if v.any():
            e = tl[:, :, k][v]
            print("tile %d: consumer wait %.0f mma %.0f | producer: requested at %.0f, landed +%.0f, published +%.0f" % (
                k, (e[:, 2] - e[:, 1]).mean(), (e[:, 4] - e[:, 3]).mean(), e[:, 6].mean(), (e[:, 7] - e[:, 6]).mean(),
                (e[:, 8] - e[:, 7]).mean()))
rdv.barrier()
rdv.close()
