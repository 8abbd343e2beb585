"""torchrun --nproc-per-node N tools/multi_gpu_check.py : csvit_allreduce_f32 (own kernel over symmetric memory) against NCCL -
values, both transports (NVSwitch multicast when mapped, two-shot P2P forced), bandwidth of a 96 MB bucket, and GradReducer
comm='symm' vs comm='nccl' on a toy model.  Rank 0 prints one JSON line per check; exit code 1 on any mismatch."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cs-vit_b200"))
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
import torch.distributed._symmetric_memory as symm
from cs_vit import ops
from cs_vit.train import GradReducer

ok = True
def report(**kw):
    if rank == 0:
        print(json.dumps(kw), flush=True)

N = 24 * 1024 * 1024                                   # 96 MB
FLAGS = ops.ALLREDUCE_FLAG_BYTES // 4
buf = symm.empty(N + FLAGS, dtype=torch.float32, device=dev)
buf.zero_()
hdl = symm.rendezvous(buf, dist.group.WORLD)
ptrs = [int(p) for p in hdl.buffer_ptrs]
mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
torch.cuda.synchronize(); dist.barrier()
report(check="rendezvous", world=world, multicast=bool(mc))

def run(n, use_mc, scale):
    ops.allreduce_f32(ptrs, [p + N * 4 for p in ptrs], mc if use_mc else 0, n, rank, world, scale)

for use_mc in ([True, False] if mc else [False]):
    for n in (4, 1024, 1000 * 4, 3 * 1024 * 1024 + 4, N):
        g = torch.Generator(device=dev).manual_seed(1234 + rank)
        src = torch.randn(n, device=dev, generator=g)
        want = src.clone(); dist.all_reduce(want); want /= world
        buf[:n].copy_(src)
        tail = buf[n:n + 8].clone() if n + 8 <= N else None
        run(n, use_mc, 1.0 / world)
        torch.cuda.synchronize()
        err = (buf[:n] - want).abs().max().item() / max(want.abs().max().item(), 1e-30)
        same = [torch.empty_like(buf[:n]) for _ in range(world)]
        dist.all_gather(same, buf[:n].contiguous())
        identical = all(torch.equal(same[0], s) for s in same)
        untouched = tail is None or torch.equal(tail, buf[n:n + 8])
        good = err < 1e-6 and identical and untouched
        ok &= good
        report(check="values", transport="multicast" if use_mc else "p2p", n=n, rel_err=err, replicas_identical=identical, tail_untouched=untouched, ok=good)
    # bandwidth, 96 MB, back to back (the kernel is its own barrier)
    for _ in range(3): run(N, use_mc, 1.0)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): run(N, use_mc, 1.0)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 20
    report(check="bandwidth", transport="multicast" if use_mc else "p2p", bytes=N * 4, us=round(us, 1), algbw_GBs=round(N * 4 / us / 1e3, 1),
           busbw_GBs=round(N * 4 / us / 1e3 * 2 * (world - 1) / world, 1))
t = torch.randn(N, device=dev)
for _ in range(3): dist.all_reduce(t)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): dist.all_reduce(t)
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / 20
report(check="bandwidth", transport="nccl all_reduce", bytes=N * 4, us=round(us, 1), algbw_GBs=round(N * 4 / us / 1e3, 1),
       busbw_GBs=round(N * 4 / us / 1e3 * 2 * (world - 1) / world, 1))

# GradReducer: symmetric transport vs NCCL on the same toy model and data
def toy():
    torch.manual_seed(3)
    return torch.nn.Sequential(torch.nn.Linear(64, 512), torch.nn.Tanh(), torch.nn.Linear(512, 512), torch.nn.Tanh(), torch.nn.Linear(512, 8)).to(dev)
grads = {}
for comm in ("symm", "nccl"):
    m = toy()
    red = GradReducer(m.parameters(), bucket_bytes=256 << 10, comm=comm)
    for step in range(3):
        g = torch.Generator(device=dev).manual_seed(50 + rank + 10 * step)
        x = torch.randn(32, 64, device=dev, generator=g)
        red.zero_grad()
        (m(x) ** 2).mean().backward()
        red.finish()
    torch.cuda.synchronize()
    grads[comm] = [p.grad.detach().clone() for p in m.parameters()]
    transport = red.transport
    report(check="reducer", comm=comm, transport=transport, buckets=len(red.bucket_summary()))
err = max(((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item() for a, b in zip(grads["symm"], grads["nccl"]))
good = err < 1e-5
ok &= good
report(check="reducer_symm_vs_nccl", max_rel_err=err, ok=good)
flag = torch.tensor([0 if ok else 1], device=dev); dist.all_reduce(flag)
torch.cuda.synchronize(); dist.barrier()
dist.destroy_process_group()
sys.exit(1 if flag.item() else 0)
