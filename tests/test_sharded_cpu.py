"""World-size-2 gloo test of the multi-GPU router's host logic (partition by owner, count
exchange, all-to-all-v, reverse path and un-permute) on the serial simulator library."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def test_row_sharding_world2_gloo():
    from hostsim import build as hb
    sim = hb.build()
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", WORLD_SIZE="2",
               SMATRIX_DIR_LOG2="8")
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "sharded_worker.py"), sim],
                              env=dict(env, RANK=str(r)), stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                              text=True) for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r} failed:\n{o}"
        assert f"rank {r} ok" in o
