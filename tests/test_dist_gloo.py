"""World-size-2 check of the by-image sharding used by bench.py / the batch API:
each rank takes its contiguous image range (no data-path collective), only the
timing reduction (max over ranks) crosses ranks.  gloo on CPU; the per-rank
compute stand-in is the oracle, since this box has no GPU."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

from oracle import wm_oracle as O  # noqa: E402
from thatsmyface_b200.pipeline import shard_ranges  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)                      # every rank sees the same batch definition
    imgs = rng.integers(0, 256, (n, 16, 24, 3), dtype=np.uint8)
    wm = rng.integers(0, 256, (2, 3), dtype=np.uint8)
    lo, hi = shard_ranges(n, world)[rank]
    mine = np.stack([O.embed_array(imgs[k], wm) for k in range(lo, hi)]) if hi > lo else np.zeros((0, 16, 24, 3), np.uint8)
    # timing reduction exactly as bench.py does it: max over ranks
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.barrier()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # result gather happens after timing, off the hot path
    parts = [None] * world
    dist.all_gather_object(parts, (lo, hi, mine))
    if rank == 0:
        full = np.concatenate([p[2] for p in sorted(parts, key=lambda p: p[0])])
        want = np.stack([O.embed_array(imgs[k], wm) for k in range(n)])
        q.put((float(t.item()), bool(np.array_equal(full, want)), [(p[0], p[1]) for p in parts]))
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [5, 8])
def test_two_ranks_shard_by_image(n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    tmax, same, ranges = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert tmax == 2.0 and same
    assert sorted(ranges) == shard_ranges(n, 2)
