"""-m gpu, needs >= 2 devices (skipped on a one-GPU box): the NCCL paths of dist.py on real hardware, one process per GPU
launched with torchrun on 127.0.0.1.  The same exchanges are covered on CPU by the world_size-2 gloo tests of
tests/test_host_logic.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _torchrun(script, n=2, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(ROOT, "tools", script)]
    r = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout)
    assert r.returncode == 0, r.stdout[-4000:]
    return r.stdout


@pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 GPUs")
def test_item_sharded_similarity_matches_single_gpu():
    _torchrun("mgpu_check.py")


@pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 GPUs")
def test_user_sharded_bpr_descends_like_single_gpu_and_replicas_agree():
    out = _torchrun("mgpu_bpr_check.py")
    assert "OK" in out and "identical across ranks: True" in out


@pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 GPUs")
def test_row_sharded_ials_matches_single_gpu():
    _torchrun("mgpu_ials_check.py")


@pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 GPUs")
def test_column_sharded_slim_matches_single_shard():
    _torchrun("mgpu_slim_check.py")


@pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 GPUs")
def test_user_sharded_ease_gram_matches_single_gpu():
    _torchrun("mgpu_ease_check.py")
