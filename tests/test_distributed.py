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


# ------------------------------------------------------------------------------------------------ finetune-step reduction
class _Toy(torch.nn.Module):
    """Stand-in with the structural features the reducer must cope with: a parameter that never gets a gradient (the
    discarded head layers, quirk Q2), a frozen parameter, and parameters of very different sizes."""

    def __init__(self):
        super().__init__()
        g = torch.Generator().manual_seed(3)
        self.a = torch.nn.Linear(16, 64)
        self.b = torch.nn.Linear(64, 64)
        self.unused = torch.nn.Linear(64, 64)
        self.frozen = torch.nn.Linear(64, 8)
        self.c = torch.nn.Linear(64, 4)
        for p in self.parameters():
            with torch.no_grad():
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)
        self.frozen.requires_grad_(False)

    def forward(self, batch):
        h = torch.tanh(self.b(torch.relu(self.a(batch["x"]))))
        return {"loss": ((self.c(h) - batch["y"]) ** 2).mean() + self.frozen(h).mean()}


def _toy_data(n):
    g = torch.Generator().manual_seed(17)
    return {"x": torch.randn(n, 16, generator=g), "y": torch.randn(n, 4, generator=g)}


def _reduce_worker(rank, world, port, q):
    import sys
    sys.path.insert(0, os.path.join(ROOT, "cs-vit_b200"))
    from cs_vit.train import GradReducer, broadcast_parameters, finetune_step, scaled_lr
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(100 + rank)            # ranks start from different weights ...
        model = _Toy()
        with torch.no_grad():
            model.a.weight.add_(rank)
        broadcast_parameters(model)               # ... and must leave with rank 0's (DDP constructor semantics, C1)
        ref = _Toy()
        assert all(torch.equal(p, r) for p, r in zip(model.parameters(), ref.parameters()))
        data = _toy_data(8)
        per = 8 // world
        mine = {k: v[rank * per:(rank + 1) * per] for k, v in data.items()}
        reducer = GradReducer(model.parameters(), bucket_bytes=8 << 10)     # small buckets -> several of them
        opt = torch.optim.SGD([p for p in model.parameters() if p.requires_grad], lr=scaled_lr(0.05, world, per))
        ropt = torch.optim.SGD([p for p in ref.parameters() if p.requires_grad], lr=scaled_lr(0.05, world, per))
        for step in range(3):                     # step 1 learns the bucket order, steps 2-3 use the overlapped hooks
            finetune_step(model, mine, opt, reducer, max_norm=1e9)
            # single-process truth: mean over ranks of the per-rank mean loss == loss over the whole batch here
            ropt.zero_grad(set_to_none=True)
            ref(data)["loss"].backward()
            for (n, p), r in zip(model.named_parameters(), ref.parameters()):
                if r.grad is None:
                    assert p.grad is None, n
                else:
                    assert torch.allclose(p.grad, r.grad, rtol=1e-5, atol=1e-7), (step, n)
            ropt.step()
            for p, r in zip(model.parameters(), ref.parameters()):
                assert torch.allclose(p, r, rtol=1e-5, atol=1e-7)
        assert len(reducer.bucket_summary()) >= 2
        assert model.unused.weight.grad is None and model.frozen.weight.grad is None
        q.put((rank, "ok"))
    finally:
        dist.destroy_process_group()


def test_grad_reducer_world2_matches_full_batch():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_reduce_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert sorted(q.get(timeout=5)[0] for _ in range(world)) == [0, 1]


class _Branchy(torch.nn.Module):
    """Which parameters receive a gradient depends on the data: ``extra`` is used only when ``batch["use_extra"]`` is set."""

    def __init__(self):
        super().__init__()
        g = torch.Generator().manual_seed(5)
        self.a = torch.nn.Linear(8, 8)
        self.extra = torch.nn.Linear(8, 8)
        self.late = torch.nn.Linear(8, 8)
        for p in self.parameters():
            with torch.no_grad():
                p.copy_(torch.randn(p.shape, generator=g) * 0.2)

    def forward(self, batch):
        h = self.a(batch["x"])
        if batch["use_extra"]:
            h = h + self.extra(batch["x"])
        if batch["use_late"]:
            h = h + self.late(batch["x"])
        loss = (h ** 2).mean()
        if batch.get("poison"):
            loss = loss * float("nan")
        return {"loss": loss}


def _divergent_worker(rank, world, port, q):
    import sys
    sys.path.insert(0, os.path.join(ROOT, "cs-vit_b200"))
    from cs_vit import train
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model = _Branchy()
        reducer = train.GradReducer(model.parameters(), bucket_bytes=1 << 10)
        opt = torch.optim.SGD(model.parameters(), lr=0.1)
        g = torch.Generator().manual_seed(40 + rank)
        # step 1: only rank 1 touches `extra`; step 2: only rank 0; step 3: `late` appears (on rank 1 only) after the buckets were built;
        # step 4: rank 0's loss is NaN -> every rank skips the batch, nobody hangs, nobody steps
        plan = [(rank == 1, False, False), (rank == 0, False, False), (False, rank == 1, False), (False, False, rank == 0)]
        for step, (ue, ul, poison) in enumerate(plan):
            before = [p.detach().clone() for p in model.parameters()]
            batch = {"x": torch.randn(4, 8, generator=g), "use_extra": ue, "use_late": ul, "poison": poison}
            train.finetune_step(model, batch, opt, reducer, max_norm=1e9)
            flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
            both = [torch.zeros_like(flat) for _ in range(world)]
            dist.all_gather(both, flat)
            assert torch.equal(both[0], both[1]), f"replicas diverged at step {step}"
            changed = any(not torch.equal(b, p) for b, p in zip(before, model.parameters()))
            assert changed == (step < 3), (step, changed)
            if step < 2:
                assert model.extra.weight.grad is not None          # the averaged gradient reaches the rank that had none
            if step == 2:
                assert model.late.weight.grad is not None
        assert train.skipped_steps == 1
        q.put((rank, "ok"))
    finally:
        dist.destroy_process_group()


def test_grad_reducer_rank_divergent_gradients_and_nan_guard():
    """ranks that differ in which parameters get gradients must still issue matching collectives and keep their
    replicas identical; a non-finite loss on one rank skips the batch on all ranks (ref:scripts/finetune.py:219-222)."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_divergent_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert sorted(q.get(timeout=5)[0] for _ in range(world)) == [0, 1]


def test_scaled_lr_rule():
    import sys
    sys.path.insert(0, os.path.join(ROOT, "cs-vit_b200"))
    from cs_vit.train import scaled_lr
    assert abs(scaled_lr(1e-4, 8, 32) - (8 * 32 / 44) ** 0.5 * 1e-4) < 1e-12      # ref:scripts/finetune.py:138-139
    assert scaled_lr(1e-4, 1, 44) == 1e-4
