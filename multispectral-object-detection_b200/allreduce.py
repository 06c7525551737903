"""The one exchange step of the reference's data-parallel training: the gradient all-reduce.

The reference wraps the two-stream model in ``DistributedDataParallel`` (``train.py:654-658``): after every backward
the gradients of all ranks are summed and divided by the world size (the loss is pre-multiplied by ``world_size`` at
``train.py:760`` to undo the averaging).  Nothing else crosses GPUs -- BatchNorm statistics stay local unless
``--sync-bn`` (``train.py:938``), and the forward shards by image pair with no collective (``shard.py``).

``GradientAllReduce`` is that step as explicit plumbing over ``torch.distributed`` (NCCL over NVLink/NVSwitch on the GPUs,
gloo in the CPU tests): parameters are packed, last layer first (the order backward produces them), into flat buckets; each
bucket is reduced with one asynchronous all-reduce on a side stream as soon as it is filled, so the reduction of the head's
gradients overlaps whatever still computes; ``finish()`` waits, averages and scatters the result back into ``p.grad``.

Status: the exchange of BASELINE config 4.  ``scripts/train_step.py`` drives it under a train step whose backward is PyTorch
autograd over the reference's modules (own backward kernels are not built) and times step / all-reduce /
overlap against ``DistributedDataParallel`` on the same box; ``tests/test_allreduce_gpu.py`` checks it against a plain
synchronous all-reduce over NCCL.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist


class GradientAllReduce:
    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_bytes: int = 25 << 20,
                 dtype: Optional[torch.dtype] = None, group=None):
        """``bucket_bytes``: DDP's default bucket is 25 MiB.  ``dtype``: wire dtype of the buckets (``None`` = the
        gradients' own dtype; ``torch.bfloat16`` halves the bytes on the wire, summation then happens in bf16)."""
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("GradientAllReduce: no trainable parameters")
        self.group = group
        self.dtype = dtype
        self.buckets: List[List[torch.nn.Parameter]] = []
        cur: List[torch.nn.Parameter] = []
        cur_bytes = 0
        for p in reversed(self.params):                       # backward produces the last layers' gradients first
            nbytes = p.numel() * (torch.empty((), dtype=dtype or p.dtype).element_size())
            if cur and cur_bytes + nbytes > bucket_bytes:
                self.buckets.append(cur)
                cur, cur_bytes = [], 0
            cur.append(p)
            cur_bytes += nbytes
        if cur:
            self.buckets.append(cur)
        self._flat: List[Optional[torch.Tensor]] = [None] * len(self.buckets)
        self._work: List = []
        self._stream = None
        self._bucket_of = {id(p): i for i, b in enumerate(self.buckets) for p in b}
        self._ready = [0] * len(self.buckets)
        self._hooks: List = []

    # ------------------------------------------------------------------ overlap with backward
    def attach(self) -> "GradientAllReduce":
        """Launch every bucket's all-reduce from inside ``backward`` as soon as the last gradient of the bucket has been
        accumulated (``register_post_accumulate_grad_hook``) -- what DDP's reducer does (train.py:654-658) -- instead of after
        the whole backward.  Call ``finish()`` after ``backward``; ``detach()`` removes the hooks."""
        if self._hooks:
            return self
        for p in self.params:
            self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))
        return self

    def detach(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks = []

    def _on_grad(self, p) -> None:
        i = self._bucket_of[id(p)]
        self._ready[i] += 1
        if self._ready[i] == len(self.buckets[i]) and self._flat[i] is None:
            self._launch(i)

    # ------------------------------------------------------------------ sizes
    @property
    def numel(self) -> int:
        return sum(p.numel() for p in self.params)

    def wire_bytes(self) -> int:
        return sum(p.numel() * torch.empty((), dtype=self.dtype or p.dtype).element_size() for p in self.params)

    # ------------------------------------------------------------------ the exchange
    def _world(self) -> int:
        return dist.get_world_size(self.group) if (dist.is_available() and dist.is_initialized()) else 1

    def _launch(self, i: int) -> None:
        bucket = self.buckets[i]
        dev = self.params[0].device
        cuda = dev.type == "cuda"
        if cuda and self._stream is None:
            self._stream = torch.cuda.Stream(dev)
        if cuda:
            self._stream.wait_stream(torch.cuda.current_stream(dev))     # the gradients' producer stream
        for p in bucket:
            if p.grad is None:
                raise RuntimeError("GradientAllReduce: a parameter has no gradient (unused parameters are not "
                                   "supported, as with DDP's default find_unused_parameters=False)")
        ctx = torch.cuda.stream(self._stream) if cuda else _Null()
        with ctx:
            wire = self.dtype or bucket[0].grad.dtype
            flat = torch.cat([p.grad.reshape(-1).to(wire) for p in bucket])
            self._flat[i] = flat
            if self._world() > 1:
                self._work.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def start(self) -> None:
        """Pack every bucket that has not been launched by a hook yet and launch its all-reduce (asynchronously; on CUDA on
        a side stream that waits for the gradients' producer stream)."""
        if self._work and not self._hooks:
            raise RuntimeError("GradientAllReduce.start(): previous exchange not finished")
        for i in range(len(self.buckets)):
            if self._flat[i] is None:
                self._launch(i)

    def finish(self) -> None:
        """Wait for the reductions, divide by the world size (DDP semantics) and write the result back to ``p.grad``."""
        if any(f is None for f in self._flat):
            self.start()                    # hook mode: buckets whose parameters got no gradient hook this step
        world = self._world()
        dev = self.params[0].device
        cuda = dev.type == "cuda"
        ctx = torch.cuda.stream(self._stream) if cuda else _Null()
        with ctx:
            # Work.wait() on NCCL orders the CURRENT stream after the collective: call it with the side stream current,
            # the stream the division and the scatter below run on (gloo's wait() simply blocks the host)
            for w in self._work:
                w.wait()
            self._work = []
            for bucket, flat in zip(self.buckets, self._flat):
                if flat is None:
                    raise RuntimeError("GradientAllReduce.finish() without start()")
                if world > 1:
                    flat = flat / world
                off = 0
                for p in bucket:
                    n = p.numel()
                    p.grad.copy_(flat[off:off + n].view_as(p.grad))
                    off += n
        if cuda:
            torch.cuda.current_stream(dev).wait_stream(self._stream)
        self._flat = [None] * len(self.buckets)
        self._ready = [0] * len(self.buckets)

    def reduce(self) -> None:
        self.start()
        self.finish()


class _Null:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


def parameter_shapes(model: torch.nn.Module) -> Sequence[torch.Size]:
    return [p.shape for p in model.parameters() if p.requires_grad]
