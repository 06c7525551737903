"""world_size-2 gloo test (CPU) of the training path's one exchange step (reference train.py:654-658, DDP gradient
averaging): bucketed asynchronous all-reduce == the plain mean of the two ranks' gradients, for fp32 and bf16 wire formats,
on the parameter set of a real (small) x3 graph."""
import importlib
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _grads(params, rank):
    g = torch.Generator().manual_seed(100 + rank)
    return [torch.randn(p.shape, generator=g) for p in params]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    cft = importlib.import_module("multispectral-object-detection_b200")
    ar = importlib.import_module("multispectral-object-detection_b200.allreduce")
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    model = cft.Model(cft.named_config("yolov5s_fusion_transformerx3_vedai"))
    params = [p for p in model.parameters() if p.requires_grad]
    expect = [(a + b) / 2 for a, b in zip(_grads(params, 0), _grads(params, 1))]
    ok = {}
    for name, dtype, tol in (("fp32", None, 1e-6), ("bf16", torch.bfloat16, 2e-2)):
        for p, g in zip(params, _grads(params, rank)):
            p.grad = g.clone()
        red = ar.GradientAllReduce(params, bucket_bytes=4 << 20, dtype=dtype)
        red.reduce()
        err = max(float((p.grad - e).abs().max()) for p, e in zip(params, expect))
        ok[name] = (err <= tol, len(red.buckets), red.numel == sum(p.numel() for p in params), red.wire_bytes())
    dist.destroy_process_group()
    q.put((rank, ok))


def test_gradient_allreduce_world2_matches_mean():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok in res:
        for name, (good, nb, all_params, wire) in ok.items():
            assert good, (rank, name)
            assert nb > 3 and all_params                             # yolov5s-x3: 44.5 M parameters in 4 MiB buckets
        assert ok["bf16"][3] * 2 == ok["fp32"][3]                    # half the bytes on the wire


def test_gradient_allreduce_single_process_is_identity():
    ar = importlib.import_module("multispectral-object-detection_b200.allreduce")
    lin = torch.nn.Linear(8, 4)
    for p in lin.parameters():
        p.grad = torch.ones_like(p) * 3
    red = ar.GradientAllReduce(lin.parameters(), bucket_bytes=16)
    assert len(red.buckets) == 2 and red.buckets[0][0] is lin.bias          # last parameter first
    red.reduce()
    assert all(bool((p.grad == 3).all()) for p in lin.parameters())


def _hook_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    ar = importlib.import_module("multispectral-object-detection_b200.allreduce")
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.ReLU(), torch.nn.Linear(32, 32), torch.nn.ReLU(),
                              torch.nn.Linear(32, 4))
    params = list(net.parameters())
    red = ar.GradientAllReduce(params, bucket_bytes=2048).attach()           # several buckets, launched from backward
    xs = [torch.randn(8, 16, generator=torch.Generator().manual_seed(7 + r)) for r in range(world)]
    grads = []
    for r in range(world):                                                   # what each rank computes on its own
        net.zero_grad()
        red.detach()
        net(xs[r]).square().mean().backward()
        grads.append([p.grad.clone() for p in params])
    expect = [sum(g[i] for g in grads) / world for i in range(len(params))]
    errs = []
    for step in range(2):                                                    # two steps: the per-step state resets
        net.zero_grad()
        red.attach()
        net(xs[rank]).square().mean().backward()
        launched_in_backward = sum(f is not None for f in red._flat)
        red.finish()
        errs.append(max(float((p.grad - e).abs().max()) for p, e in zip(params, expect)))
    dist.destroy_process_group()
    q.put((rank, errs, launched_in_backward, len(red.buckets)))


def test_gradient_allreduce_hooks_launch_buckets_during_backward():
    """attach(): every bucket's all-reduce is launched by the post-accumulate-grad hook of its last parameter (the overlap
    DDP's reducer provides, train.py:654-658); the averaged gradients equal the plain mean."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_hook_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, errs, launched, nb in res:
        assert max(errs) <= 1e-6, (rank, errs)
        assert nb >= 3 and launched == nb
