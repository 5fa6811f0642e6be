#!/usr/bin/env python
"""Headline benchmark: walker-steps/s of the red-blue walker update.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one full ensemble step (every walker proposed once = the P
half-steps of ``RedBlueMove.propose``).  Default workload: BASELINE.json's
headline -- 65 536 walkers, 128-D correlated Gaussian (dense inverse
covariance), StretchMove(a=2), randomised split, fp64 -- on each GPU
(weak scaling: N GPUs step N x 65 536 walkers sharded by row block).

Prints ONE JSON line (rank 0).  ``value`` is device-timed (CUDA events on the
engine's stream, max over ranks) with the state resident in HBM; ``e2e`` is
the same metric through the public API ``EnsembleSampler.run_mcmc`` with host
buffers, copies inside the timed region.  ``--impl reference`` times the CPU
oracle port (numpy restatement of the reference, pinned bit-exact against it)
on the host cores.
"""

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "walker-steps/sec (nwalkers x iters / s)"
MODEL_SEED, INIT_SEED, SAMPLER_SEED = 20240, 20241, 0x656D636565B200


# --------------------------------------------------------------------------
# synthetic workloads (SURVEY 8d / DESIGN.md); plain numpy parameter generation
# --------------------------------------------------------------------------
def make_workload(name, nwalkers, ndim):
    rng_m = np.random.default_rng(MODEL_SEED)
    rng_p = np.random.default_rng(INIT_SEED)
    w = {"name": name, "nwalkers": nwalkers, "ndim": ndim}
    if name == "gauss_dense":
        v = rng_m.standard_normal((ndim + 1, ndim))  # random_cov(ndim, dof=1), document/plots/oned.py:21-25
        cov = (v.T @ v) / (ndim + 1)
        icov = np.linalg.inv(cov)
        w["icov"] = 0.5 * (icov + icov.T)
        w["p0"] = rng_p.standard_normal((nwalkers, ndim))
        w["moves"] = "stretch"
    elif name == "gauss_iso":
        w["p0"] = rng_p.standard_normal((nwalkers, ndim))
        w["moves"] = "stretch"
    elif name == "ring":
        w["radius"], w["sigma"] = 5.0, 0.5
        w["p0"] = rng_p.standard_normal((nwalkers, ndim)) * (5.0 / np.sqrt(ndim))
        w["moves"] = "stretch"
    elif name == "rosenbrock":
        w["p0"] = 1.0 + 0.1 * rng_p.standard_normal((nwalkers, ndim))
        w["moves"] = "de+snooker"
    else:
        raise ValueError(name)
    return w


def flops_bytes_per_walker_step(w):
    """Algorithmic work of one walker-step: fp64 bytes with the row
    write counted unconditionally, and flops of proposal + log-prob.  For the
    dense Gaussian the engine evaluates -0.5 |L^T x|^2 with A = L L^T, whose
    algorithmic cost is D(D+1) + 2D flops (SURVEY 8d quotes 2D^2 + 3D for the
    unfactored x^T A x; DESIGN.md explains the choice)."""
    D = w["ndim"]
    if w["moves"] == "stretch":
        nbytes, prop = 24 * D + 24, 3 * D
    else:  # 0.8 DE + 0.2 snooker
        nbytes = 0.8 * (32 * D + 24) + 0.2 * (40 * D + 24)
        prop = 0.8 * 3 * D + 0.2 * 10 * D
    lp = {"gauss_dense": D * (D + 1) + 2 * D, "gauss_iso": 2 * D, "ring": 2 * D + 6, "rosenbrock": 9 * (D - 1)}[w["name"]]
    return prop + lp, nbytes


# --------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------
class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons during the timed region
    (B200_PROFILING.md clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL,
            )
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(smax)), reasons=sorted(reasons), samples=len(sm))
        return out


def kernel_traffic(kernel):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture
    (profiles/traffic.json, written by scripts/summarize_ncu.py), or None."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        return t.get(kernel)
    except Exception:
        return None


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def host_threads():
    try:
        from threadpoolctl import threadpool_info

        n = [p.get("num_threads", 1) for p in threadpool_info() if p.get("user_api") == "blas"]
        return int(max(n)) if n else 1
    except Exception:
        return 1


# --------------------------------------------------------------------------
# CPU arm: the oracle port (bench.py may execute oracle/ only here)
# --------------------------------------------------------------------------
def oracle_sampler(w, seed):
    from oracle import redblue as rb
    from oracle import targets as T

    D = w["ndim"]
    target = {
        "gauss_dense": lambda: T.GaussDense(w["icov"]),
        "gauss_iso": lambda: T.GaussIso(D),
        "ring": lambda: T.Ring(D, w["radius"], w["sigma"]),
        "rosenbrock": lambda: T.Rosenbrock(D),
    }[w["name"]]()
    moves = [(rb.Stretch(), 1.0)] if w["moves"] == "stretch" else [(rb.DE(), 0.8), (rb.Snooker(), 0.2)]
    o = rb.OracleSampler(w["nwalkers"], D, target, moves, seed=seed)
    o.set_state(w["p0"])
    return o


def time_oracle(w, steps, warmup):
    o = oracle_sampler(w, SAMPLER_SEED)
    o.run(warmup)
    t0 = time.perf_counter()
    o.run(steps)
    dt = time.perf_counter() - t0
    return w["nwalkers"] * steps / dt, dt


def run_reference(args, dist):
    """``--impl reference``: the reference's CPU algorithm (oracle port:
    vectorised numpy restatement, pinned bit-exact against dfm/emcee@8ab6c0f by
    tests/golden) on this box's host cores.  Rank 0 only."""
    if dist.rank != 0:
        return
    w = make_workload(args.workload, args.nwalkers, args.ndim)
    value, dt = time_oracle(w, args.steps, max(1, min(args.warmup, 3)))
    line = {
        "impl": "reference",
        "metric": METRIC, "value": value, "unit": "walker-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": max(1, min(args.warmup, 3)), "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args, w, 1),
        "cpu_baseline": {
            "value": value, "unit": "walker-steps/s", "cores": host_threads(), "kind": "port",
            "sample": "%d full-ensemble steps of the %dx%d %s workload (numpy oracle port; BLAS threads=%d, rest single-threaded)"
            % (args.steps, w["nwalkers"], w["ndim"], w["name"], host_threads()),
        },
        "e2e": {"value": value, "unit": "walker-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args, w, world, n_local=None):
    n_local = w["nwalkers"] if n_local is None else n_local
    return {
        "workload": "%s %dx%d per GPU, %s, fp64, randomize_split=True, %s"
        % (w["name"], n_local, w["ndim"],
           "StretchMove(a=2)" if w["moves"] == "stretch" else "0.8 DEMove + 0.2 DESnookerMove",
           "ensemble = %d walkers sharded by row block over %d GPUs" % (n_local * world, world) if world > 1 else "1 GPU"),
        "nwalkers_total": n_local * world,
        "ndim": w["ndim"],
        "l2": args.l2_note,
        "comm": args.comm if world > 1 else "none",
    }


# --------------------------------------------------------------------------
# the B200 arm
# --------------------------------------------------------------------------
def run_b200(args, dist):
    import emcee_b200
    from emcee_b200 import _lib, models, moves

    world = dist.world
    if args.scaling == "weak":
        n_local, n_total = args.nwalkers, args.nwalkers * world
    else:
        n_total = args.nwalkers
        n_local = n_total // world
    w = make_workload(args.workload, n_total, args.ndim)
    D = w["ndim"]
    model = {
        "gauss_dense": lambda: models.GaussianDense(w["icov"]),
        "gauss_iso": lambda: models.GaussianIso(),
        "ring": lambda: models.Ring(w["radius"], w["sigma"]),
        "rosenbrock": lambda: models.Rosenbrock(),
    }[w["name"]]()
    mv = moves.StretchMove() if w["moves"] == "stretch" else [(moves.DEMove(), 0.8), (moves.DESnookerMove(), 0.2)]

    sampler = emcee_b200.EnsembleSampler(n_total, D, model, moves=mv, seed=SAMPLER_SEED, device=dist.local_rank,
                                        pinned_results=True)
    eng = sampler._engine
    from emcee_b200 import dist as ebdist

    ebdist.attach(eng, dist, args.comm)
    if world > 1 and args.l2_flush:
        # N >= 2 GPUs of weak scaling: the ensemble (>= 128 MiB) no longer fits the 126 MB L2, so the
        # timing rule is met by size; the per-step flush brackets are a single-GPU device
        args.l2_flush = False
        args.l2_note = ("no flush: the ensemble (%d MiB) is larger than L2 (126 MB), as are the split tables"
                        % (n_total * D * 8 >> 20))
    eng.set_option("l2_flush", 1 if args.l2_flush else 0)
    if args.dmma_group > 0:
        eng.set_option("dmma_group", args.dmma_group)
    if args.no_tma_rows:
        eng.set_option("tma_rows", 0)
    if args.tma_rows:
        eng.set_option("tma_rows", 1)
    if args.no_stagger:
        eng.set_option("dmma_stagger", 0)
    if args.no_pdl:
        eng.set_option("pdl", 0)
    sched = sampler._schedule()

    # ---- device-resident throughput (`value`) ------------------------------
    eng.set_state(w["p0"])
    eng.step(sched, args.warmup, want_accepted=False)
    clocks = ClockSampler(dist.local_rank)
    dist.barrier()
    clocks.start()
    t0 = time.perf_counter()
    eng.step(sched, args.steps, want_accepted=False)  # synchronous at return
    wall = time.perf_counter() - t0
    dist.barrier()
    ck = clocks.stop()
    ck["source"] = "timed region"
    ms, launches = eng.last_step_timing()
    ms = dist.max(ms)
    wall = dist.max(wall)
    if ck["samples"] < 3 and world == 1:
        # the timed region is shorter than nvidia-smi can resolve: sample the same workload for ~0.7 s
        # more (same repeat count on every rank: the steps contain cross-rank barriers)
        reps = max(1, min(200, int(0.7 / max(ms * 1e-3, 1e-4))))
        probe = ClockSampler(dist.local_rank)
        probe.start()
        for _ in range(reps):
            eng.step(sched, args.steps, want_accepted=False)
        ck2 = probe.stop()
        if ck2["samples"] > ck["samples"]:
            ck = ck2
            ck["source"] = "same workload repeated for ~0.7 s right after the timed region (region too short to sample)"
    value = n_total * args.steps / (ms * 1e-3)
    kernel = eng.last_kernel_name()

    # ---- end to end through the public API with host buffers (`e2e`) --------
    # The call a user makes: run_mcmc(p0, K, store=False) -- H2D of the initial
    # state from pinned host memory, K steps, D2H of the final state.
    p0_pinned = _lib.pinned_empty((n_total, D))
    p0_pinned[...] = w["p0"]
    sampler.run_mcmc(p0_pinned, max(3, args.warmup // 4), store=False, skip_initial_state_check=True)
    dist.barrier()
    t0 = time.perf_counter()
    last = sampler.run_mcmc(p0_pinned, args.steps, store=False, skip_initial_state_check=True)
    e2e_wall = dist.max(time.perf_counter() - t0)
    e2e_value = n_total * args.steps / e2e_wall
    h2d = n_total * D * 8 / args.steps
    d2h = (n_total * D * 8 + n_total * 8) / args.steps
    assert np.all(np.isfinite(last.log_prob))

    if dist.rank != 0:
        return

    flops, nbytes = flops_bytes_per_walker_step(w)
    peaks = measured_peaks()
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "B200_PROFILING.md fallback 6650 GB/s (of fallback)"
    per_gpu_rate = n_local * args.steps / (ms * 1e-3)  # walker-steps/s of one GPU's kernels
    ach_gbs = per_gpu_rate * nbytes / 1e9
    ach_tf = per_gpu_rate * flops / 1e12
    fp64_peak, fp64_src = None, ("fp64 issue-rate peak measured on this GPU by eb_microbench (max of DMMA m8n8k4 and DFMA); "
                                 "MEASURED_PEAKS.json has no fp64 entry")
    if args.no_microbench:
        fp64_peak, fp64_src = 37.0, "recorded eb_microbench DMMA m8n8k4 peak (profiles/r01_fp64_microbench.txt)"
    else:
        try:
            fp64_peak = max(_lib.microbench(1, 16), _lib.microbench(0, 32))
        except Exception:
            pass
    roofline_hbm = {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                    "traffic": None, "peak_source": hbm_src,
                    "note": "algorithmic bytes/walker-step = %g (SURVEY 8d) over the whole timed region (launch gaps included)" % nbytes}
    if w["name"] == "gauss_dense" and fp64_peak:
        roofline = {"bound": "tensor", "achieved": ach_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach_tf / fp64_peak,
                    "traffic": kernel_traffic("dense_dmma"),
                    "peak_source": fp64_src,
                    "note": "algorithmic flops/walker-step = %g (triangular D(D+1)+2D log-prob + 3D proposal; "
                            "the unfactored 2D^2+3D form would read 2x higher) over the whole timed region" % flops}
    else:
        roofline = roofline_hbm

    # ---- CPU baseline beside it (bounded sample, rank 0 only, N=1 only) -----
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        wc = make_workload(args.workload, min(n_total, args.cpu_nwalkers), D)
        steps_c = args.cpu_steps
        v, dt = time_oracle(wc, steps_c, 1)
        cpu = {"value": v, "unit": "walker-steps/s", "cores": host_threads(), "kind": "port",
               "sample": "%d steps of %dx%d %s on the oracle port (numpy; %.1f s)" % (steps_c, wc["nwalkers"], D, w["name"], dt)}

    line = {
        "metric": METRIC, "value": value, "unit": "walker-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, w, world, n_local),
        "clocks": {"sm_mhz": ck["sm_mhz"], "sm_max_mhz": ck["sm_max_mhz"], "reasons": ck["reasons"], "samples": ck["samples"],
                   "source": ck["source"]},
        "e2e": {"value": e2e_value, "unit": "walker-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "call": "EnsembleSampler.run_mcmc(p0_pinned_host, %d, store=False): H2D initial state + %d steps + D2H final state, wall clock"
                % (args.steps, args.steps)},
        "gpu_launches": launches,
        "kernel": kernel,
        "wall_ms_per_step": 1e3 * wall / args.steps,
        "roofline": roofline,
        "roofline_hbm": roofline_hbm,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="gauss_dense", choices=["gauss_dense", "gauss_iso", "ring", "rosenbrock"])
    ap.add_argument("--nwalkers", type=int, default=65536, help="walkers per GPU (weak) or in total (strong)")
    ap.add_argument("--ndim", type=int, default=128)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--comm", default="p2p", choices=["allgather", "p2p"],
                    help="multi-GPU exchange: NVLink peer-memory pull (default) or one ncclAllGather per split")
    ap.add_argument("--no-l2-flush", dest="l2_flush", action="store_false")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-microbench", action="store_true",
                    help="do not launch the fp64 peak micro-benchmarks (for ncu launch lists); use the recorded peak")
    ap.add_argument("--tma-rows", action="store_true", help="force the tma_rows kernel on")
    ap.add_argument("--no-tma-rows", action="store_true", help="HBM-bound models: use the generic kernel instead of tma_rows")
    ap.add_argument("--no-stagger", action="store_true", help="dense_dmma: all pairs request their first tile at once")
    ap.add_argument("--no-pdl", action="store_true", help="dense_dmma: plain stream-ordered launches instead of programmatic dependent launches")
    ap.add_argument("--dmma-group", type=int, default=0, help="half-steps per persistent dense_dmma launch (0: library default)")
    ap.add_argument("--cpu-steps", type=int, default=40)
    ap.add_argument("--cpu-nwalkers", type=int, default=65536)
    args = ap.parse_args()
    args.warmup = max(3, args.warmup)
    args.l2_note = (
        "L2 flushed between timed steps (256 MiB device memset, outside the per-step CUDA-event brackets)"
        if args.l2_flush else
        "no flush: the ensemble (N*D*8 B) stays L2-resident across steps, as it does in a real run"
    )
    from emcee_b200.dist import Rendezvous

    dist = Rendezvous()
    if dist.world != args.gpus and dist.world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, dist.world))
    try:
        if args.impl == "reference":
            run_reference(args, dist)
        else:
            run_b200(args, dist)
    finally:
        dist.close()


if __name__ == "__main__":
    main()
