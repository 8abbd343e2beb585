"""CPU suite, part 3: the N>1 host logic on world_size-2 gloo (no GPU): batch sharding covers the batch exactly
once and keeps clips whole; the single-collective eval gather returns what the reference's five collectives
(ref:scripts/eval.py:53-82, 289-292) would, in rank order, including ragged last batches."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make_batch(n, T=2):
    g = torch.Generator().manual_seed(n)
    return {
        "patches": torch.rand(n, T, 3, 8, 8, generator=g),
        "square_bboxes": torch.rand(n, T, 4, generator=g),
        "timestamp": torch.arange(T).float()[None].repeat(n, 1),
        "imgs_path": [[f"/data/seq{i}/f{t}.jpg" for t in range(T)] for i in range(n)],
        "flip": [bool(i % 2) for i in range(n)],
        "scalar": 3,
    }


def _results(n, seed):
    g = torch.Generator().manual_seed(seed)
    return {"joint_cam_gt": torch.randn(n, 21, 3, generator=g), "joint_cam_pred": torch.randn(n, 21, 3, generator=g),
            "joint_reproj_gt": torch.randn(n, 21, 2, generator=g), "joint_reproj_pred": torch.randn(n, 21, 2, generator=g)}


def _worker(rank, world, port, n_total, q):
    import sys
    sys.path.insert(0, os.path.join(ROOT, "cs-vit_b200"))
    from cs_vit.distributed import gather_eval_results, shard_batch, shard_range
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        batch = _make_batch(n_total)
        mine = shard_batch(batch, rank, world)
        lo, hi = shard_range(n_total, rank, world)
        assert mine["patches"].shape[0] == hi - lo and mine["patches"].shape[1] == 2      # clips stay whole
        assert torch.equal(mine["patches"], batch["patches"][lo:hi]) and mine["flip"] == batch["flip"][lo:hi]
        assert mine["scalar"] == 3
        full = _results(n_total, 99)
        local = {k: v[lo:hi] for k, v in full.items()}
        paths = [p[-1] for p in batch["imgs_path"][lo:hi]]
        max_local = -(-n_total // world)
        got = gather_eval_results(local, paths, max_local=max_local)
        if rank == 0:
            merged, all_paths = got
            for k, v in full.items():
                assert torch.equal(merged[k], v), k                                       # bit-exact, rank order
            assert all_paths == [p[-1] for p in batch["imgs_path"]]
        else:
            assert got is None
        q.put((rank, "ok"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [8, 7, 1])
def test_shard_and_gather_world2(n_total):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert sorted(q.get(timeout=5)[0] for _ in range(world)) == [0, 1]


def test_shard_range_partitions():
    import sys
    sys.path.insert(0, os.path.join(ROOT, "cs-vit_b200"))
    from cs_vit.distributed import shard_range
    for n in (0, 1, 7, 256, 257):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_single_process_gather_without_init():
    import sys
    sys.path.insert(0, os.path.join(ROOT, "cs-vit_b200"))
    from cs_vit.distributed import gather_eval_results
    r = _results(3, 1)
    merged, paths = gather_eval_results(r, ["a", "b/c.jpg", "d"])
    assert paths == ["a", "b/c.jpg", "d"] and torch.equal(merged["joint_reproj_pred"], r["joint_reproj_pred"])
