"""Multi-GPU parity on real devices: the three sharding modes under torchrun + NCCL against the
single-process oracle.  Skipped below two devices (the driver's one-GPU box); run with `gpurun --gpus 2`."""
import json
import math
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def W_seconds(mjd, start):
    return (mjd[0] - start[0]) * 86400.0 + (mjd[1] - start[1]) + (mjd[2] - start[2])


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["time", "channel", "subband"])
def test_sharded_pipeline_matches_oracle(case, tmp_path):
    n = _ngpu()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (found %d)" % n)
    world = 2 if n < 4 or case == "time" else 4
    out = tmp_path / ("%s.json" % case)
    port = 29650 + {"time": 0, "channel": 1, "subband": 2}[case]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_worker.py"),
           "--case", case, "--out", str(out)]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    res = json.loads(out.read_text())
    assert res["backend"] == "nccl" and res["world"] == world
    assert res["hits_equal"], res
    assert res["ndat_total"] == res["ndat_expected"], res
    assert res["err"] <= 1e-5, res


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["time", "channel"])
def test_native_multi_host_combines_like_phase_series(mode):
    """libb200multi.so (include/b200multi.h): ONE process, one worker thread + stream + pipeline per GPU, blocks fed
    from host memory, NCCL reduce (time shards) or per-device copy-out (channel shards) at the sub-integration
    boundary, attributes merged with the PhaseSeries::combine rules -- against the single-process oracle
    (tests/mgpu_native.py, run in its own process)."""
    n = _ngpu()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (found %d)" % n)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "mgpu_native.py"), mode], cwd=ROOT,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert json.loads(r.stdout.strip().splitlines()[-1])["ok"]
