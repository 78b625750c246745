"""Multi-GPU parity on real NCCL (needs >= 2 CUDA devices; skipped otherwise): every rank computes the loss of its image
shard through dist.sharded_focal_loss - kernel-written shard statistics, NCCL all-gather, g3d_combine_shard_stats, gradient
scales verified on the device - and compares the global losses and its local gradients with the single-GPU loss of the
full batch (uneven shards, an empty image).  The world_size-2 gloo test of tests/test_host_logic.py covers the same host
functions on the CPU."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("peer_exchange", ["1", "0"])
@pytest.mark.parametrize("world", [2, 4])
def test_sharded_loss_parity_multi_gpu(world, peer_exchange):
    """peer_exchange = 1: the statistics travel as stores into the peers' memory (g3d_exchange_shard_stats); 0: NCCL
    all-gather + g3d_combine_shard_stats.  Same losses, same gradients."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} CUDA devices, found {torch.cuda.device_count()}")
    port = 29600 + (os.getpid() % 300) + world + 10 * int(peer_exchange)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "dist_worker.py")]
    env = dict(os.environ, G3D_PEER_EXCHANGE=peer_exchange)
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0 and "check_dist OK" in r.stdout, r.stdout[-3000:]
    if peer_exchange == "0":
        assert "via NCCL all-gather" in r.stdout
    print(r.stdout[-400:])
