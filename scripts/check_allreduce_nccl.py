#!/usr/bin/env python
"""NCCL check of GradientAllReduce: bucketed asynchronous all-reduce on a side stream (reduce() and the
hook-driven attach() mode) == a plain synchronous all-reduce of every gradient, with work queued on the producer stream right
before and right after the exchange (the stream-ordering bug a host-blocking gloo wait() cannot show).  Run with
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port P scripts/check_allreduce_nccl.py
Exit code 0 = every rank agrees; prints one line on rank 0."""
import importlib
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ar = importlib.import_module("multispectral-object-detection_b200.allreduce")
    torch.manual_seed(0)
    net = torch.nn.Sequential(*[torch.nn.Linear(1024, 1024) for _ in range(12)]).to(dev)       # 12.6 M parameters
    params = list(net.parameters())
    red = ar.GradientAllReduce(params, bucket_bytes=4 << 20)
    worst = 0.0
    for it in range(5):
        g = torch.Generator(device=dev).manual_seed(100 * it + rank)
        x = torch.randn(256, 1024, device=dev, generator=g)
        # reference: synchronous all-reduce of the gradients of the same forward / backward
        net.zero_grad()
        red.detach()
        net(x).square().mean().backward()
        expect = []
        for p in params:
            e = p.grad.clone()
            dist.all_reduce(e)
            expect.append(e / world)
        for mode in ("reduce", "hooks"):
            net.zero_grad()
            if mode == "hooks":
                red.attach()
            net(x).square().mean().backward()
            if mode == "reduce":
                red.detach()
                for p in params:                       # late writes on the producer stream, right before the exchange
                    p.grad.mul_(1.0)
                red.reduce()
            else:
                red.finish()
            got = [p.grad.clone() for p in params]     # reads on the producer stream, right after the exchange
            torch.cuda.synchronize()
            worst = max(worst, max(float((a - b).abs().max() / (b.abs().max() + 1e-12)) for a, b in zip(got, expect)))
    red.detach()
    t = torch.tensor([worst], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = float(t.item()) <= 1e-5
    if rank == 0:
        print(f"check_allreduce_nccl: world {world}, buckets {len(red.buckets)}, worst rel diff {float(t.item()):.2e} -> {'ok' if ok else 'MISMATCH'}")
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
