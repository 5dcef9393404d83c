"""N > 1 on real GPUs (skipped on a 1-GPU box): torchrun + NCCL Monte-Carlo check and a 2-rank bench line."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(n, script, *args, port=29533):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), script, *args]
    return subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_monte_carlo_counters_independent_of_gpu_count():
    r = _torchrun(2, os.path.join(ROOT, "tests", "dist_mc_check.py"))
    assert r.returncode == 0 and "DIST_MC_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_bench_two_ranks_prints_one_json_line():
    r = _torchrun(2, os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "2", "--warmup", "3", "--frames", "8192",
                  port=29534)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["n_gpus"] == 2 and d["scaling"] == "weak" and d["config"]["global_frames"] == 2 * 8192
    assert d["avg_iterations"] == 10.0 and d["e2e"]["matches_device_path"]
