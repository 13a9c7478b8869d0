"""The subtree-sharded path over REAL NCCL: one process per GPU (torchrun), world size 2 (and 4 when the box has them).
Skipped on a single-GPU box; the same logic with the ranks simulated in one process is tests/test_sharded.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize('world', [2, 4])
def test_device_shards_over_nccl(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip('needs {} GPUs'.format(world))
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world),
           '--master-addr', '127.0.0.1', '--master-port', str(29640 + world), os.path.join(HERE, 'nccl_shard_worker.py')]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert 'NCCL_SHARD_OK world={}'.format(world) in res.stdout
