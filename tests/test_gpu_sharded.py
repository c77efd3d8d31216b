"""Multi-GPU parity (BASELINE config 5 shape at test scale): rows hash-partitioned over the GPUs of
one box, ops routed by NCCL all-to-all, every rank updating only its own shard — bit-exact against
the reference fed the whole stream.  Needs >= 2 GPUs (`gpurun --gpus 2`)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("p2p", [1, 0], ids=["peer-memory", "nccl-a2a"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_row_sharding_nccl(world, p2p):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(29600 + world + 10 * p2p),
               WORLD_SIZE=str(world), SMX_P2P=str(p2p))
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "sharded_gpu_worker.py")],
                              env=dict(env, RANK=str(r), LOCAL_RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(world)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r} failed:\n{o[-3000:]}"
        assert f"rank {r} ok" in o
