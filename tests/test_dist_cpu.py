"""world_size-2 gloo test of the N>1 host logic (candidate sharding + score gather/selection)."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import _pkg
    d = _pkg.import_sub("dist")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_cand = 5  # ragged: 3 + 2
    mine = d.shard_range(n_cand, rank, world)
    all_scores = np.array([-3.0, -1.5, -2.0, -0.25, -4.0], dtype=np.float32)
    winner, owner, local, scores = d.gather_select(all_scores[list(mine)], [10 + c for c in mine])
    q.put((rank, list(mine), winner, owner, local, scores.tolist()))
    dist.destroy_process_group()


def test_shard_and_select_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 1, 2] and res[1][1] == [3, 4]
    for r in res:
        assert r[2] == 3 and r[3] == 1 and r[4] == 0  # candidate 3 wins, lives on rank 1 as local 0
        assert np.allclose(r[5], [-3.0, -1.5, -2.0, -0.25, -4.0])


def test_shard_range_covers_everything():
    sys.path.insert(0, ROOT)
    import _pkg
    d = _pkg.import_sub("dist")
    for n in (1, 7, 64, 256):
        for w in (1, 2, 4, 8):
            got = [c for r in range(w) for c in d.shard_range(n, r, w)]
            assert got == list(range(n))
            sizes = [len(d.shard_range(n, r, w)) for r in range(w)]
            assert max(sizes) - min(sizes) <= 1
