"""2-GPU data-parallel parity (skipped on single-GPU boxes): see tests/dp_worker.py."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_step_matches_single_rank():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29631",
                          os.path.join(ROOT, "tests", "dp_worker.py")], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "dp parity ok" in res.stdout


def test_two_rank_attribution_matches_single_rank():
    """the attribution pass shards by image; two ranks must reproduce one process's averages / node IE."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29633",
                          os.path.join(ROOT, "tests", "dp_ie_worker.py")], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "dp ie parity ok" in res.stdout


def test_two_rank_cuda_graphed_pipeline_matches_eager():
    """ModelPipeline(data_parallel=True, cuda_graph=True): the captured batch (frozen forward, step_grads, the peer-memory
    exchange with its device-side counter, step_apply, comparison) replayed on two ranks equals the eager run."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, SVB_COMM_TIMEOUT_S="30")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29635",
                          os.path.join(ROOT, "tests", "dp_graph_worker.py")], capture_output=True, text=True, timeout=600,
                         env=env)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "dp graph parity ok" in res.stdout
