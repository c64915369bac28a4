"""Host logic of the N>1 path on CPU: the band partition used by bench.py under
torchrun (rank r renders bands r, r+N, ...) covers every line exactly once for
world sizes 1..8 and AA factors 1..3, and a world_size-2 gloo job agrees on the
aggregate the way bench.py computes it (MAX of times, SUM of iterations)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rank_lines(user_height, aa, rank, world):
    return [l for b in range(rank, user_height, world) for l in range(b * aa, (b + 1) * aa)]


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("aa", [1, 2, 3])
def test_band_partition_covers_image_once(world, aa):
    user_h = 37
    seen = np.zeros(user_h * aa, dtype=int)
    for r in range(world):
        for l in rank_lines(user_h, aa, r, world):
            seen[l] += 1
    assert (seen == 1).all()


WORKER = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.environ["MDZ_ROOT"]); sys.path.insert(0, os.path.join(os.environ["MDZ_ROOT"], "tests"))
import portpath
from views import config2
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
view = config2(64, 36, 300)
full = portpath.port_render(view, 1)                 # CPU oracle stands in for the device
mine = [l for b in range(rank, view.user_height, world) for l in range(b, b + 1)]
part = np.where(full[mine] > 0, full[mine], view.depth).astype(np.int64).sum()
it = torch.tensor([int(part)], dtype=torch.int64)
t = torch.tensor([1.0 + rank], dtype=torch.float64)
dist.all_reduce(it, op=dist.ReduceOp.SUM)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
whole = int(np.where(full > 0, full, view.depth).astype(np.int64).sum())
assert int(it[0]) == whole, (int(it[0]), whole)
assert float(t[0]) == float(world)
if rank == 0:
    print("OK", whole)
dist.destroy_process_group()
'''


def test_gloo_world2_aggregate(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MDZ_ROOT=ROOT)
    import socket
    with socket.socket() as sk:                      # a free rendezvous port: a fixed one may sit in TIME_WAIT
        sk.bind(("127.0.0.1", 0))
        port = str(sk.getsockname()[1])
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", port, str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
