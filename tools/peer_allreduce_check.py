"""torchrun --nproc-per-node N tools/peer_allreduce_check.py : one-shot NVLink all-reduce vs NCCL
(values and latency) for the trainer's 6,603-float exchange."""
import json
import os
import sys
import time
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from doper_b200.distributed import PeerAllReducer  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 6603
red = PeerAllReducer(n, dev)
ok = True
g = torch.Generator(device=dev); g.manual_seed(100 + rank)
for step in range(200):
    x = torch.randn(n, device=dev, generator=g)
    red.slot().copy_(x)
    got = red.all_reduce().clone()
    ref = x.clone()
    dist.all_reduce(ref)
    if not torch.allclose(got, ref, rtol=1e-6, atol=1e-6):
        ok = False
        print(f"rank {rank} step {step}: max err {(got - ref).abs().max().item()}", flush=True)
        break
# every rank must hold the same bits
chk = got.clone()
dist.broadcast(chk, 0)
same_bits = bool(torch.equal(chk, got))
status = int(red.status.item())


def timeit(fn, reps=300):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    dist.barrier(device_ids=[local])
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return 1e3 * a.elapsed_time(b) / reps


x = torch.randn(n, device=dev)
buf = x.clone()


def peer():
    red.slot().copy_(x)
    red.all_reduce()


def nccl():
    buf.copy_(x)
    dist.all_reduce(buf)


t_peer, t_nccl = timeit(peer), timeit(nccl)
res = torch.tensor([t_peer, t_nccl], device=dev, dtype=torch.float64)
dist.all_reduce(res, op=dist.ReduceOp.MAX)
if rank == 0:
    line = {"world": world, "floats": n, "values_match_nccl": ok, "same_bits_on_every_rank": same_bits,
            "status": status, "peer_us": round(float(res[0]), 2), "nccl_us": round(float(res[1]), 2)}
    print(json.dumps(line), flush=True)
    Path("gpurun_out").mkdir(exist_ok=True)
    Path(f"gpurun_out/peer_allreduce_n{world}.json").write_text(json.dumps(line))
red.close()
dist.destroy_process_group()
