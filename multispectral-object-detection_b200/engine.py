"""Serving-side executor for the two-stream forward: CUDA streams and graphs, no tracing compiler.

``ForwardEngine`` wraps a ``Model`` for one fixed input geometry (batch, H, W, uint8 wire format
``[B,6,H,W]`` = RGB||IR, reference ``utils/datasets.py:1272-1281`` / ``train.py:715-717``):

* the whole 47-layer forward (~350 kernel launches) is captured once per buffer slot into a CUDA graph
  (launch-bound at small batch) and replayed;
* two (or more) device input/output slots + two copy streams (H2D, D2H): the host->device copy of batch i+1 and the
  device->host copy of detections i-1 overlap the compute of batch i;
* ``concurrent=True`` gives every slot its own compute stream so that consecutive batches overlap on the GPU as well.
  Measured (profiles/r01_timeline.md, v15): device-resident throughput +0–2 % (run-to-run noise is 1.5 %), end-to-end
  throughput −2–3 % (the copy pipeline of a slot no longer hides behind exactly one other compute), so the default is
  one compute stream, slots serialised;
* ``infer(host_u8)`` is the blocking call a user makes; ``submit()/collect()`` expose the pipeline;
* ``nms={...}`` (arguments of ``nms.nms_batched``) appends the batched NMS kernel to the captured graph, so that the
  per-step device->host copy is ``[B, max_det, 6]`` detections + counts (0.23 MB at batch 32) instead of the raw
  ``z [B, 25200, no]`` (25.8 MB) -- the step ``detect_twostream.py:83-86`` performs on the host side of the reference.

The reference has no counterpart (it launches eager PyTorch ops on the default stream, ``test.py:106-119``).
"""
from __future__ import annotations

from typing import List, Optional

import torch

from . import nms as _nms
from ._lib import CftError, launch_count


class ForwardEngine:
    def __init__(self, model, batch: int, height: int, width: int, device=None, slots: int = 2, use_graph: bool = True,
                 nms: Optional[dict] = None, concurrent: bool = False):
        self.model = model.eval()
        self.nms_kw = dict(nms) if nms is not None else None
        self.device = torch.device(device) if device is not None else next(model.parameters()).device
        if self.device.type != "cuda":
            raise CftError("ForwardEngine needs a CUDA device (no CPU fallback)")
        self.shape = (batch, 6, height, width)
        self.slots = slots
        self.use_graph = use_graph
        self.concurrent = bool(concurrent) and slots > 1
        n_streams = slots if self.concurrent else 1
        self.computes = [torch.cuda.Stream(self.device) for _ in range(n_streams)]
        self.compute = self.computes[0]               # slot s runs on computes[s] (all on computes[0] when not concurrent)
        self.h2d = torch.cuda.Stream(self.device)     # separate copy streams: a D2H queued behind a compute event
        self.d2h = torch.cuda.Stream(self.device)     # must not block the next batch's H2D
        self.x_dev = [torch.zeros(self.shape, dtype=torch.uint8, device=self.device) for _ in range(slots)]
        self.z_dev: List[Optional[torch.Tensor]] = [None] * slots
        self.graphs: List[Optional[torch.cuda.CUDAGraph]] = [None] * slots
        self.ev_in = [torch.cuda.Event() for _ in range(slots)]
        self.ev_done = [torch.cuda.Event() for _ in range(slots)]
        self.ev_free = [torch.cuda.Event() for _ in range(slots)]
        self.z_host: List[Optional[torch.Tensor]] = [None] * slots
        self.cnt_dev: List[Optional[torch.Tensor]] = [None] * slots      # with nms: z_dev/z_host hold the detections
        self.cnt_host: List[Optional[torch.Tensor]] = [None] * slots
        self._nms_ws: Optional[List[torch.Tensor]] = None
        self.launches_per_forward = 0
        self._next = 0
        self._pending: List[int] = []
        self._build()

    def stream_of(self, slot: int) -> "torch.cuda.Stream":
        return self.computes[slot] if self.concurrent else self.computes[0]

    # ------------------------------------------------------------------ setup
    def _forward(self, s: int):
        x = self.x_dev[s]
        z, _ = self.model(x[:, :3], x[:, 3:])
        if self.nms_kw is None:
            return z
        if self._nms_ws is None:                       # sized once, outside any capture (the warm-up pass comes first);
            b, rows, no = z.shape                      # one workspace per slot: slots run concurrently
            ml = bool(self.nms_kw.get("multi_label", False))
            self._nms_ws = [torch.empty((b * rows * (no - 5 if ml else 1),), dtype=torch.int64, device=self.device)
                            for _ in range(self.slots)]
        det, cnt = _nms.nms_batched(z, workspace=self._nms_ws[s], **self.nms_kw)
        self.cnt_dev[s] = cnt
        return det

    def _build(self):
        torch.cuda.synchronize(self.device)
        with torch.cuda.stream(self.compute), torch.no_grad():
            for _ in range(2):                      # warm-up: packs weights, sizes the allocator pools
                self._forward(0)
            n0 = launch_count()
            self._forward(0)
            self.launches_per_forward = launch_count() - n0
        self.compute.synchronize()
        for s in range(self.slots):
            st = self.stream_of(s)
            if self.use_graph:
                g = torch.cuda.CUDAGraph()
                with torch.no_grad(), torch.cuda.graph(g, stream=st):
                    self.z_dev[s] = self._forward(s)
                self.graphs[s] = g
            else:
                with torch.cuda.stream(st), torch.no_grad():
                    self.z_dev[s] = self._forward(s)
                st.synchronize()
            self.z_host[s] = torch.empty(self.z_dev[s].shape, dtype=self.z_dev[s].dtype).pin_memory()
            if self.nms_kw is not None:
                self.cnt_host[s] = torch.empty(self.cnt_dev[s].shape, dtype=torch.int32).pin_memory()
            self.ev_free[s].record(st)
        for st in self.computes:
            st.synchronize()

    # ------------------------------------------------------------------ device-resident replay (kernel-only timing)
    def run_resident(self, slot: int = 0):
        """One forward over whatever is in the slot's device input buffer, on the slot's compute stream."""
        with torch.cuda.stream(self.stream_of(slot)):
            if self.use_graph:
                self.graphs[slot].replay()
            else:
                with torch.no_grad():
                    self.z_dev[slot] = self._forward(slot)
        return self.z_dev[slot]

    # ------------------------------------------------------------------ pipelined host->device->host path
    def submit(self, host_u8: torch.Tensor) -> int:
        """Queue one batch (uint8 [B,6,H,W], ideally pinned). Returns the slot to ``collect``."""
        if tuple(host_u8.shape) != self.shape or host_u8.dtype != torch.uint8:
            raise CftError(f"submit: expected uint8 {self.shape}, got {host_u8.dtype} {tuple(host_u8.shape)}")
        s = self._next
        self._next = (self._next + 1) % self.slots
        if s in self._pending:
            raise CftError("submit: all slots in flight; collect() first")
        with torch.cuda.stream(self.h2d):
            self.h2d.wait_event(self.ev_free[s])            # slot's previous result has left the device
            self.x_dev[s].copy_(host_u8, non_blocking=True)
            self.ev_in[s].record(self.h2d)
        st = self.stream_of(s)
        with torch.cuda.stream(st):
            st.wait_event(self.ev_in[s])
            if self.use_graph:
                self.graphs[s].replay()
            else:
                with torch.no_grad():
                    self.z_dev[s] = self._forward(s)
            self.ev_done[s].record(st)
        with torch.cuda.stream(self.d2h):
            self.d2h.wait_event(self.ev_done[s])
            self.z_host[s].copy_(self.z_dev[s], non_blocking=True)
            if self.nms_kw is not None:
                self.cnt_host[s].copy_(self.cnt_dev[s], non_blocking=True)
            self.ev_free[s].record(self.d2h)
        self._pending.append(s)
        return s

    def collect(self, slot: Optional[int] = None):
        """Wait for the oldest (or the given) submitted batch; returns its detections z on the host (pinned), or with
        ``nms=...`` the pair (det [B, max_det, 6], counts [B]) -- rows [0, counts[b]) of image b are valid."""
        if not self._pending:
            raise CftError("collect: nothing submitted")
        s = self._pending.pop(0) if slot is None else self._pending.pop(self._pending.index(slot))
        self.ev_free[s].synchronize()
        if self.nms_kw is not None:
            return self.z_host[s], self.cnt_host[s]
        return self.z_host[s]

    def infer(self, host_u8: torch.Tensor):
        """Blocking convenience call: host uint8 batch in, host fp32 detections ``z [B, rows, no]`` out
        (``(det, counts)`` with ``nms=...``)."""
        return self.collect(self.submit(host_u8))

    def drain(self):
        while self._pending:
            self.collect()
        for st in self.computes:
            st.synchronize()
        self.h2d.synchronize()
        self.d2h.synchronize()
