#!/usr/bin/env python
"""BASELINE config 4's exchange step, timed alone: the gradient all-reduce of the yolov5l-x3 parameter set (206.3 M
parameters) over NCCL, one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           scripts/allreduce_row.py [--dtype bf16|fp32] [--bucket-mb 25] [--steps 10]

Synthetic gradients of the real parameter shapes (the backward kernels of this path are not built);
CUDA events around `GradientAllReduce.reduce()`, max over ranks; rank 0 prints one JSON line with the time per exchange and
the bus bandwidth 2 (N-1)/N x bytes / time."""
import argparse
import importlib
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--bucket-mb", type=int, default=25)
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pkg = importlib.import_module("multispectral-object-detection_b200")
    ar = importlib.import_module("multispectral-object-detection_b200.allreduce")
    model = pkg.Model(pkg.named_config("yolov5l_fusion_transformerx3_FLIR_aligned")).to(dev)
    params = [p for p in model.parameters() if p.requires_grad]
    for p in params:
        p.grad = torch.randn_like(p)
    red = ar.GradientAllReduce(params, bucket_bytes=args.bucket_mb << 20,
                               dtype=torch.bfloat16 if args.dtype == "bf16" else None)
    for _ in range(3):
        red.reduce()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        red.reduce()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = float(t.item())
        wire = red.wire_bytes()
        print(json.dumps({"what": "gradient all-reduce, yolov5l-x3 parameter set", "n_gpus": world, "params": red.numel,
                          "wire_dtype": args.dtype, "wire_mb": round(wire / 1e6, 1), "buckets": len(red.buckets),
                          "ms_per_exchange": round(ms, 3),
                          "bus_gbs": round(2 * (world - 1) / max(world, 1) * wire / (ms / 1e3) / 1e9, 1) if world > 1 else None}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
