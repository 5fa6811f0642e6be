#!/usr/bin/env python
"""Peer-memory read bandwidth for the engine's access patterns: torchrun ... scripts/probe_nvlink.py"""
import sys
sys.path.insert(0, ".")
import numpy as np
import emcee_b200
from emcee_b200 import dist, models

rdv = dist.Rendezvous("gloo")
for D in (128, 32):
    N = 65536 * 128 // D
    s = emcee_b200.EnsembleSampler(N, D, models.GaussianIso(), seed=1, device=rdv.local_rank)
    eng = s._engine
    dist.attach(eng, rdv, "p2p")
    eng.set_state(np.random.default_rng(0).standard_normal((N, D)))
    names = {0: "stream 16B loads", 1: "random rows, 16B loads", 2: "random rows, TMA bulk"}
    for solo in (True, False):
        for what in (0, 1, 2):
            rdv.barrier()
            res = {}
            for peer in sorted({rdv.rank, (rdv.rank + 1) % rdv.world}):
                if solo and rdv.rank != 0:
                    continue
                res[peer] = eng.comm_probe(peer, what)
            rdv.barrier()
            if rdv.rank == 0:
                print("D=%3d row=%4dB %-26s %s: %s" % (D, D * 8, names[what], "rank0 only" if solo else "all ranks at once",
                      ", ".join("%s %.0f GB/s" % ("local" if p == 0 else "peer%d" % p, v) for p, v in res.items())), flush=True)
    eng.close()
rdv.close()
