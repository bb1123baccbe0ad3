"""Multi-GPU paths on real devices (skipped on a single-GPU box): the copy-engine panel push with the flag-polling DGEMM
kernel (multigpu.TiledGemm, p2p_push) and the bulk-mode partitioned SGEMM/ZGEMM/DSYRK/DTRSM (partitioned.py) must be
bit-identical to the single-GPU routines."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


def _torchrun(n, script, *args, timeout=900):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", "29577", os.path.join(ROOT, script)] + list(args)
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    return [json.loads(l) for l in out.stdout.splitlines() if l.startswith("{")]


def test_tiled_dgemm_p2p_push_two_gpus():
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    lines = _torchrun(2, "bench.py", "--gpus", "2", "--steps", "2", "--warmup", "1", "--size", "4096", "--verify")
    assert lines and lines[-1]["n_gpus"] == 2 and lines[-1]["verified"]["max_abs_diff_vs_1gpu"] == 0.0
    assert "copy engines" in lines[-1]["config"]["parallelism"]


def test_partitioned_routines_two_gpus():
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    lines = _torchrun(2, "tools/partitioned_check.py", "sgemm", "zgemm", "dsyrk", "dtrsm")
    assert len(lines) == 4
    for l in lines:
        assert l["max_abs_diff_vs_1gpu"] == 0.0, l
