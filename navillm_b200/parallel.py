"""Data-parallel wrapper of the path: the one-line replacement for the reference's DDP wrap (tools/optims.py:52-54).

    # from torch.nn.parallel import DistributedDataParallel as DDP
    from navillm_b200.parallel import DistributedDataParallel as DDP
    model = DDP(model, device_ids=[device_id], find_unused_parameters=True)

Everything else of the reference stays as it is: ``isinstance(model, torch.nn.parallel.DistributedDataParallel)`` and
``model.no_sync`` in the rollout loop (tasks/agents/mp3d_agent.py:661-667), ``model.module.lang_model.cls_token`` (:817),
``save_checkpoint`` unwrapping ``model.module`` (tools/optims.py:66-67), ``clip_grad_norm_(model.parameters())`` and the
optimizer step with NO explicit reduce in train.py:86-89.

What is different from torch's DDP is the mechanism.  NavModel keeps its gradients in two flat buffers (bf16 LM + heads,
fp32 encoder / embeddings) that its hand-written backward accumulates into natively, so there are no per-parameter
autograd hooks and no buckets to rebuild.  ``GradSync`` reproduces DDP's *semantics* on top of that:

* a forward outside ``no_sync()`` with grad enabled ARMS the exchange, a forward inside ``no_sync()`` disarms it
  (DDP decides at forward time, from the last forward before the backward);
* the first custom backward node of an armed pass queues an end-of-backward callback on the autograd engine
  (``Variable._execution_engine.queue_callback`` -- the hook DDP itself finalises with): when the whole backward of
  the pass (head -> LM -> fusion glue -> panorama encoder) has run, every flat buffer is all-reduced (AVG) exactly once;
* while the LM backward is running, the gradient slice of each finished group of decoder layers is all-reduced
  asynchronously so NCCL overlaps the remaining wgrad GEMMs; the end-of-backward callback waits for those handles and
  reduces only what they did not cover;
* an unarmed backward (inside ``no_sync``) issues NO collective, so ranks whose rollouts have different lengths issue
  the same number of collectives as under the reference's DDP: one exchange per armed pass.

The class derives from ``torch.nn.parallel.DistributedDataParallel`` only so that the reference's ``isinstance`` test
holds; torch's reducer is never constructed.
"""
from __future__ import annotations

import contextlib
from typing import Callable, List, Optional

import torch
import torch.nn as nn


def _dist():
    import torch.distributed as dist
    return dist if (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1) else None


class GradSync:
    """State machine of the gradient exchange of ONE model replica (shared by NavModel and its language model)."""

    def __init__(self):
        self.armed = False            # the last grad-enabled forward ran outside no_sync()
        self.queued = False           # the end-of-backward callback of the running pass is queued
        self.overlap = True           # reduce finished layer groups while the backward is still running
        self.chunk_layers = 4
        self.flats: Callable[[], list] = lambda: []          # -> [(FlatParams, tail_offset or None), ...]
        self._pending: List[tuple] = []
        self._lm_layers_reduced = False
        self.stats = {"collectives": 0, "exchanges": 0, "async_slices": 0}

    # ---- forward side -------------------------------------------------------------------------------------------
    def on_forward(self, sync: bool) -> None:
        if torch.is_grad_enabled():
            self.armed = bool(sync) and _dist() is not None

    # ---- backward side ------------------------------------------------------------------------------------------
    def backward_begins(self) -> None:
        """Called by the first custom backward node of a pass (action head, or the LM in the loss modes)."""
        if self.armed and not self.queued:
            self.queued = True
            torch.autograd.Variable._execution_engine.queue_callback(self._end_of_backward)

    def layer_hook(self, flat, starts: List[int], n_layers: int) -> Optional[Callable[[int], None]]:
        """Per-layer callback for LlamaCore.backward (layers finish from n_layers-1 down to 0): all-reduces the flat
        gradient slice [starts[l], starts[l + chunk]) of a finished group asynchronously.  None unless this pass is armed."""
        dist = _dist()
        if not (self.armed and self.queued and self.overlap and dist is not None):
            return None
        chunk = max(1, self.chunk_layers)
        avg = dist.ReduceOp.AVG if dist.get_backend() == "nccl" else dist.ReduceOp.SUM
        ws = dist.get_world_size()
        self._lm_layers_reduced = True

        def done(l: int) -> None:
            if l % chunk != 0:
                return
            sl = flat.flat_grad[starts[l]:starts[min(l + chunk, n_layers)]]
            h = dist.all_reduce(sl, op=avg, async_op=True)
            self._pending.append((h, sl if avg == dist.ReduceOp.SUM else None, ws))
            self.stats["collectives"] += 1
            self.stats["async_slices"] += 1
        return done

    def _drain(self) -> None:
        for h, sl, ws in self._pending:
            h.wait()
            if sl is not None:
                sl.div_(ws)
        self._pending = []

    def _end_of_backward(self) -> None:
        try:
            self.exchange(covered_lm_layers=self._lm_layers_reduced)
        finally:
            self.armed = self.queued = False
            self._lm_layers_reduced = False

    # ---- the exchange itself ------------------------------------------------------------------------------------
    def exchange(self, covered_lm_layers: bool = False) -> int:
        """All-reduce (average) every flat gradient buffer: ONE collective per buffer, minus the LM layer
        slices the overlapped reductions of this pass already covered.  Returns the number of collectives issued here."""
        dist = _dist()
        self._drain()
        if dist is None:
            return 0
        avg = dist.ReduceOp.AVG if dist.get_backend() == "nccl" else dist.ReduceOp.SUM
        ws, n = dist.get_world_size(), 0
        for flat, tail in self.flats():
            if flat is None:
                continue
            buf = flat.flat_grad[tail:] if (covered_lm_layers and tail is not None) else flat.flat_grad
            dist.all_reduce(buf, op=avg)
            if avg == dist.ReduceOp.SUM:
                buf.div_(ws)
            n += 1
        self.stats["collectives"] += n
        self.stats["exchanges"] += 1
        return n


class DistributedDataParallel(torch.nn.parallel.DistributedDataParallel):
    """Drop-in for ``torch.nn.parallel.DistributedDataParallel(model, device_ids=[...], find_unused_parameters=True)``
    around a ``navillm_b200.nav_model.NavModel`` (see the module docstring).  Replicas must start from identical
    parameters (same seed / same checkpoint), as the reference's do; ``broadcast_parameters()`` enforces it explicitly."""

    def __init__(self, module: nn.Module, device_ids=None, output_device=None, find_unused_parameters: bool = False,
                 broadcast_parameters: bool = True, **unused):
        nn.Module.__init__(self)                      # deliberately NOT torch DDP's __init__: no reducer, no buckets
        if not hasattr(module, "grad_sync"):
            raise TypeError("navillm_b200.parallel.DistributedDataParallel wraps a navillm_b200 NavModel "
                            f"(an object with a .grad_sync state), got {type(module).__name__}")
        self.module = module
        self.device_ids = list(device_ids) if device_ids is not None else None
        self.output_device = output_device
        self.find_unused_parameters = find_unused_parameters
        self.require_backward_grad_sync = True
        if broadcast_parameters:
            self.broadcast_parameters()

    def broadcast_parameters(self, src: int = 0) -> None:
        """DDP's construction-time parameter broadcast: one broadcast per flat weight buffer once they exist
        (``NavModel._ensure``), else per parameter."""
        dist = _dist()
        if dist is None:
            return
        flats = [f for f, _ in self.module.grad_sync.flats() if f is not None] if hasattr(self.module, "_flat_buffers_ready") \
            and self.module._flat_buffers_ready() else []
        if flats:
            for f in flats:
                dist.broadcast(f.flat, src=src)
        else:
            with torch.no_grad():
                for p in self.module.parameters():
                    dist.broadcast(p.data, src=src)

    def forward(self, *args, **kwargs):
        self.module.grad_sync.on_forward(self.require_backward_grad_sync)
        return self.module(*args, **kwargs)

    @contextlib.contextmanager
    def no_sync(self):
        """torch DDP's context (used by tasks/agents/mp3d_agent.py:661-667): backwards of forwards that ran inside
        accumulate locally; the first pass outside exchanges the accumulated gradients."""
        old = self.require_backward_grad_sync
        self.require_backward_grad_sync = False
        try:
            yield
        finally:
            self.require_backward_grad_sync = old

    # nn.Module plumbing that torch DDP overrides with reducer-dependent code
    def train(self, mode: bool = True):
        nn.Module.train(self, mode)
        return self

    def __getstate__(self):
        return self.__dict__

    def __setstate__(self, state):
        self.__dict__.update(state)
