"""CPU: host-side logic — weight generator, sharding, world_size-2 gloo gather."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

from fac_fake_b200 import weights as W
from fac_fake_b200.sharding import gather_scores, shard_by_crops, shard_range


def test_state_dict_is_deterministic_and_complete():
    a = W.make_state_dict(0, "bn")
    b = W.make_state_dict(0, "bn")
    assert list(a.keys()) == list(b.keys())
    assert all(torch.equal(a[k], b[k]) for k in a)
    assert len(a) == 193
    n_params = sum(v.numel() for k, v in a.items() if "running" not in k and "num_batches" not in k)
    assert n_params == 89_022_274 - 0 or n_params > 88_000_000
    c = W.make_state_dict(1, "bn")
    assert not torch.equal(a["cls_token"], c["cls_token"])
    x = W.synthetic_crops(3, seed=4)
    assert x.shape == (3, 224, 224, 3) and x.dtype == torch.uint8
    assert torch.equal(x, W.synthetic_crops(3, seed=4))


@pytest.mark.parametrize("n,world", [(8192, 8), (10, 4), (3, 8), (0, 2), (17, 1)])
def test_shard_range_partitions(n, world):
    spans = [shard_range(n, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
    sizes = [b - a for a, b in spans]
    assert max(sizes) - min(sizes) <= 1


def test_shard_by_crops_balances():
    counts = [30, 5, 5, 30, 30, 1, 1, 1, 29, 15]
    spans = shard_by_crops(counts, 3)
    assert spans[0][0] == 0 and spans[-1][1] == len(counts)
    assert all(spans[i][1] == spans[i + 1][0] for i in range(2))
    loads = [sum(counts[a:b]) for a, b in spans]
    assert max(loads) <= 1.6 * (sum(counts) / 3)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_videos, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n_videos, rank, world)
    local = torch.arange(lo, hi, dtype=torch.float32) * 0.5     # "score" of video v is v/2
    full = gather_scores(local, n_videos, rank, world)
    q.put((rank, full.tolist()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_videos", [7, 8])
def test_gather_scores_world2_gloo(n_videos):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_videos, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [v * 0.5 for v in range(n_videos)]
    for _, full in got:
        assert full == want


def test_blazeface_host_nms_matches_oracle():
    """The product's blending NMS (host side of BlazeFaceEngine) against the oracle restatement, random boxes."""
    from oracle import blazeface_oracle as B
    from fac_fake_b200.blazeface import BlazeFaceEngine
    eng = BlazeFaceEngine.__new__(BlazeFaceEngine)           # host logic only: no library / device needed
    eng.min_suppression_threshold = 0.3
    g = torch.Generator().manual_seed(0)
    for trial in range(100):
        k = int(torch.randint(0, 12, (1,), generator=g))
        c = torch.rand(k, 2, generator=g) * 0.6
        s = torch.rand(k, 2, generator=g) * 0.3 + 0.02
        det = torch.zeros(k, 17)
        det[:, 0:2], det[:, 2:4] = c, c + s
        det[:, 4:16] = torch.rand(k, 12, generator=g)
        det[:, 16] = torch.rand(k, generator=g) * 0.25 + 0.75
        a, b = B.weighted_nms(det), eng._weighted_non_max_suppression(det)
        assert len(a) == len(b), trial
        for x, y in zip(a, b):
            assert torch.allclose(x, y, atol=1e-6), trial
    assert eng.nms([torch.zeros((0, 17))])[0].shape == (0, 17)
