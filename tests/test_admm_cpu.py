"""CPU (gloo, world_size 2): the collective pose consensus (admm.PoseConsensus / ConsensusManager)
against a single-process restatement of the reference's master process
(admm_trainer.py:124-179) and dual update (consensus.py:40-76) on the same tile payloads."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, load_pkg


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def make_tiles(num_camera, n_tiles, seed):
    """Tile payloads as TILE.commit publishes them (tile.py:489-491)."""
    g = torch.Generator().manual_seed(seed)
    tiles = []
    for t in range(n_tiles):
        n = int(torch.randint(3, num_camera, (1,), generator=g))
        idx = torch.randperm(num_camera, generator=g)[:n]
        tiles.append((0.05 * torch.randn(n, 6, generator=g), idx, 0.5 + torch.rand(n, generator=g)))
    return tiles


def master_reference(tiles, num_camera, prev):
    """admm_trainer.py:137-179, literally."""
    temp = torch.zeros(num_camera, 6)
    count = torch.zeros(num_camera, dtype=torch.int32)
    weight = torch.zeros(num_camera)
    for pose, idx, conf in tiles:
        count[idx] += 1
        weight[idx] += conf
        temp[idx] += conf[..., None] * pose
    overlap = torch.where(count >= 2)[0]
    weight[weight == 0] = 1
    temp /= weight[..., None]
    dual = torch.mean(torch.abs(prev - temp))
    primal = sum(torch.mean(torch.abs(pose - temp[idx])) for pose, idx, _ in tiles) / len(tiles)
    outs = []
    for pose, idx, _ in tiles:
        l = idx.tolist()
        outs.append({"shared_poses": temp[idx], "overlap_idxs": torch.tensor([l.index(i) for i in l if i in overlap], dtype=torch.long)})
    return temp, outs, primal, dual


def _worker(rank, world, port, num_camera, n_tiles, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    load_pkg()
    from admm import PoseConsensus
    tiles = make_tiles(num_camera, n_tiles, 0)
    mine = tiles[rank::world]                          # round-robin tile -> rank, as admm_trainer.py:74-83
    pc = PoseConsensus(num_camera, "cpu")
    res = []
    for rnd in range(2):                               # two rounds: the dual residual uses the previous z
        if rnd == 1:
            mine = [(p * 0.5, i, c) for p, i, c in mine]
        out = pc.exchange(mine)
        # plain numpy through the queue (tensor hand-over by file descriptor races with worker exit)
        res.append((pc.shared_poses.numpy().copy(), [(o["shared_poses"].numpy().copy(), o["overlap_idxs"].numpy().copy()) for o in out],
                    float(pc.primal_residual), float(pc.dual_residual)))
    q.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_consensus_two_ranks_matches_master_process():
    num_camera, n_tiles, world = 40, 5, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, num_camera, n_tiles, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=150) for _ in range(world))
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    tiles = make_tiles(num_camera, n_tiles, 0)
    prev = torch.zeros(num_camera, 6)
    for rnd in range(2):
        if rnd == 1:
            tiles = [(p * 0.5, i, c) for p, i, c in tiles]
        z, outs, primal, dual = master_reference(tiles, num_camera, prev)
        prev = z
        for rank in range(world):
            zr, per_tile, pr, du = got[rank][rnd]
            zr = torch.from_numpy(zr)
            per_tile = [(torch.from_numpy(a), torch.from_numpy(b)) for a, b in per_tile]
            assert torch.allclose(zr, z, atol=1e-6)
            assert abs(pr - float(primal)) < 1e-6 and abs(du - float(dual)) < 1e-6
            for (sp, ov), want in zip(per_tile, outs[rank::world]):
                assert torch.allclose(sp, want["shared_poses"], atol=1e-6)
                assert torch.equal(ov, want["overlap_idxs"])


def test_consensus_manager_dual_update_and_penalty():
    load_pkg()
    from admm import ConsensusManager
    g = torch.Generator().manual_seed(1)
    x = torch.nn.Parameter(0.1 * torch.randn(7, 6, generator=g))
    cm = ConsensusManager(x, rho=100.0)
    assert cm() is None                                   # no overlap cameras yet (consensus.py:78-82)
    z = 0.1 * torch.randn(7, 6, generator=g)
    cm.update(z, torch.tensor([1, 4, 5]))
    u = 1.5 * (x.detach() - z)                            # consensus.py:44-46
    assert torch.allclose(cm.delta_se3, u)
    want = torch.mean(100.0 * ((x - z + u) ** 2)[torch.tensor([False, True, False, False, True, True, False])])
    loss = cm()
    assert torch.allclose(loss, want)
    loss.backward()
    assert x.grad is not None and float(x.grad[0].abs().sum()) == 0.0 and float(x.grad[1].abs().sum()) > 0
    ck = cm.export_check_point()
    cm2 = ConsensusManager(x, rho=1.0)
    cm2.load_check_point(ck)
    assert torch.allclose(cm2(), want)
