"""CPU (gloo, world_size 2): admm.DepthExchange -- the shared_depth exchange of tile.py:432-475 / admm_trainer.py:31-32 as
one padded all_gather -- against the reference's semantics restated in one process: a list indexed by the global camera id
that every tile writes its maps into (the higher tile index stored last) and every tile reads its own cameras from."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, load_pkg

H, W, NCAM = 6, 8, 20


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def tiles(seed=0):
    """(tile index, ids, maps) for 5 tiles; one tile writes nothing, cameras 3 and 7 are written by two tiles."""
    g = torch.Generator().manual_seed(seed)
    spec = {0: [1, 3, 7], 1: [], 2: [3, 9], 3: [7, 12, 13, 14], 4: [19]}
    return [(t, torch.tensor(ids, dtype=torch.long), torch.rand(len(ids), H, W, generator=g) + t) for t, ids in spec.items()]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    load_pkg()
    from admm import DepthExchange
    mine = tiles()[rank::world]
    want = torch.tensor([3, 7, 8, 19, 1] if rank == 0 else [12, 3, 0], dtype=torch.long)
    have, maps = DepthExchange(H, W, "cpu").exchange(mine, want)
    q.put((rank, have.numpy().copy(), maps.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_depth_exchange_two_ranks_matches_shared_list():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = {r: (torch.from_numpy(h), torch.from_numpy(m)) for r, h, m in (q.get(timeout=150) for _ in range(world))}
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    shared = [None] * NCAM                         # admm_trainer.py:117-118: shared_depth, one slot per camera
    for t, ids, maps in tiles():                   # tiles store in index order: the higher tile index wins
        for i, m in zip(ids.tolist(), maps):
            shared[i] = m
    for rank, want in ((0, [3, 7, 8, 19, 1]), (1, [12, 3, 0])):
        have, maps = got[rank]
        assert have.tolist() == [shared[i] is not None for i in want]
        for k, i in enumerate(want):
            if shared[i] is not None:
                assert torch.equal(maps[k], shared[i]), (rank, i)
            else:
                assert float(maps[k].abs().max()) == 0.0


def test_depth_exchange_single_process():
    load_pkg()
    from admm import DepthExchange
    have, maps = DepthExchange(H, W, "cpu").exchange(tiles(), torch.tensor([7, 2]))
    assert have.tolist() == [True, False]
    assert torch.equal(maps[0], tiles()[3][2][0])
    have, maps = DepthExchange(H, W, "cpu").exchange([], torch.tensor([1]))
    assert have.tolist() == [False]
