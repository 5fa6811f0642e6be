"""Multi-GPU parity under pytest: spawns one process per GPU (world = min(device_count, 2), or
EB_MG_WORLD) running tests/multigpu_check.py, which steps row-block sharded ensembles in both
exchange modes and compares them with the single-process oracle (bit-exact for the stretch move).
Skipped on a box with a single GPU."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _device_count():
    from emcee_b200 import _lib

    return _lib.device_count()


@pytest.mark.gpu
def test_two_gpus_reproduce_the_oracle(tmp_path):
    ndev = _device_count()
    if ndev < 2:
        pytest.skip("needs >= 2 GPUs (have %d)" % ndev)
    world = int(os.environ.get("EB_MG_WORLD", "2"))
    world = max(2, min(world, ndev))
    procs, logs = [], []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world),
                   EB_RDV_FILE=str(tmp_path / "rdv"), PYTHONUNBUFFERED="1")
        log = open(tmp_path / ("rank%d.log" % rank), "w")
        logs.append(log)
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "multigpu_check.py")],
                                      env=env, stdout=log, stderr=subprocess.STDOUT, cwd=ROOT))
    rcs = []
    for p in procs:
        try:
            rcs.append(p.wait(timeout=900))
        except subprocess.TimeoutExpired:
            p.kill()
            rcs.append(-9)
    for log in logs:
        log.close()
    out = "".join("--- rank %d ---\n%s" % (r, open(tmp_path / ("rank%d.log" % r)).read()) for r in range(world))
    keep = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(keep):
        with open(os.path.join(keep, "multigpu_check_world%d.log" % world), "w") as f:
            f.write(out)
    assert all(rc == 0 for rc in rcs), out[-6000:]
    assert "ALL MULTI-GPU CHECKS PASSED" in out
