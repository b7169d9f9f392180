"""N > 1 host logic on CPU with the gloo backend (world_size 2 and 3): sharding, noise partition, gather, min/max."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["PYTHONPATH"] = ROOT + os.pathsep + os.environ.get("PYTHONPATH", "")  # for the spawned workers

from consistencytta_b200 import distributed as D  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_global, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        prompts = ["p%d" % i for i in range(n_global)]
        mine = D.shard(prompts)
        noise = D.sharded_noise(n_global, 77, world, rank, shape=(2, 3))
        assert noise.shape[0] == len(mine)
        # fake per-rank "waveforms": row i of the global batch is filled with i
        a, b = D.shard_bounds(n_global, world, rank)
        wav = torch.arange(a, b, dtype=torch.float32)[:, None].repeat(1, 5) / 10 - 0.3
        mm = torch.stack([wav.min(), wav.max()]) if wav.numel() else torch.tensor([float("inf"), float("-inf")])
        gmm = D.global_minmax(mm)
        i16 = (wav * 100).to(torch.int16)
        full = D.gather_waveforms(i16, n_global)
        # gather to rank 0 only (two submits: both staging slots), unequal shards included
        sizes = [D.shard_bounds(n_global, world, r)[1] - D.shard_bounds(n_global, world, r)[0] for r in range(world)]
        gat = D.WaveformGatherer(sizes, (5,), "cpu")
        gat.submit(i16 + 1)
        gat.submit(i16)
        root = gat.result()
        assert (root is None) == (rank != 0)
        q.put((rank, mine, noise.tolist(), gmm.tolist(), full.tolist(), root.tolist() if root is not None else None))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_global", [(2, 8), (3, 7), (2, 1)])
def test_shard_gather_gloo(world, n_global):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_global, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # shards are a contiguous partition of the prompts in order
    assert sum((r[1] for r in res), []) == ["p%d" % i for i in range(n_global)]
    # concatenated per-rank noise == single-process noise
    g = torch.Generator().manual_seed(77)
    assert torch.equal(torch.cat([torch.tensor(r[2]).reshape(-1, 2, 3) for r in res]), torch.randn(n_global, 2, 3, generator=g))
    ref_wav = torch.arange(n_global, dtype=torch.float32)[:, None].repeat(1, 5) / 10 - 0.3
    for r in res:
        assert torch.allclose(torch.tensor(r[3]), torch.stack([ref_wav.min(), ref_wav.max()]))
        assert torch.equal(torch.tensor(r[4], dtype=torch.int16).reshape(n_global, 5), (ref_wav * 100).to(torch.int16))
    assert torch.equal(torch.tensor(res[0][5], dtype=torch.int16).reshape(n_global, 5), (ref_wav * 100).to(torch.int16))
    assert all(r[5] is None for r in res[1:])


def test_shard_bounds_balanced():
    for n in (0, 1, 7, 64, 1024):
        for w in (1, 2, 4, 8):
            b = [D.shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [y - x for x, y in b]
            assert max(sizes) - min(sizes) <= 1
