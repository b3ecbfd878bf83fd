"""World-size-2 gloo tests of the multi-process plumbing (frames shard with no data-path
collective; the streak DB is broadcast once)."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch.distributed as dist
from rain_rendering_b200 import dist as rdist, synth
rank, world, local = rdist.init_process_group("gloo")
assert world == 2
db = synth.make_streak_db(0)
tex = rdist.broadcast_streak_db_host(db.textures if rank == 0 else None, src=0)
assert len(tex) == 50 and all(np.array_equal(a, b) for a, b in zip(tex, db.textures))
# frame sharding: disjoint, covering, balanced
n = 7481
a, b = rdist.shard_range(n, rank, world)
import torch
t = torch.tensor([a, b])
out = [torch.zeros(2, dtype=torch.long) for _ in range(world)]
dist.all_gather(out, t)
spans = sorted((int(o[0]), int(o[1])) for o in out)
assert spans[0][0] == 0 and spans[-1][1] == n and spans[0][1] == spans[1][0]
assert abs((spans[0][1] - spans[0][0]) - (spans[1][1] - spans[1][0])) <= 1
dist.barrier()
open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "rank" + str(rank) + ".ok"), "w").write("ok")
""" % ROOT


def test_gloo_world_size_2_db_broadcast_and_sharding(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert (tmp_path / "rank0.ok").exists() and (tmp_path / "rank1.ok").exists(), out.stdout[-2000:] + out.stderr[-2000:]


def test_shard_range_properties():
    from rain_rendering_b200.dist import shard_range
    for n in (0, 1, 7, 64, 7481):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
