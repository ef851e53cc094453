"""Batch-parallel plumbing: one process per GPU, ONE all-reduce per optimiser step over a flat fp32 buffer of the
mapper gradients (M_v + active M_o, ~1.1 MB) — SURVEY.md 8e.  Replaces the DDP wrapper that accelerate puts around the
text encoder (reference training/coach.py:97-99,214), which also fixes the reference's dict-held object mappers
escaping DDP (reference models/net_clip_text_embedding.py:25-32).  Works with the nccl (GPU) and gloo (CPU tests)
backends of torch.distributed."""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


class FlatGradAllReducer:
    """Packs the .grad of the given parameters into one persistent flat fp32 buffer, all-reduces it (mean) and
    scatters the result back."""

    def __init__(self, params: Iterable[torch.nn.Parameter], device=None):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = device if device is not None else (self.params[0].device if self.params else "cpu")
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.views, off = [], 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()

    @property
    def world(self) -> int:
        return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1

    def broadcast_parameters_(self, src: int = 0) -> None:
        """DDP's construction-time step: every rank starts from rank `src`'s parameter values (module init draws from the
        process-local default generator, which torch seeds differently per process).  One broadcast of the flat buffer."""
        if not self.params or self.world == 1:
            return
        with torch.no_grad():
            for p, v in zip(self.params, self.views):
                v.copy_(p)
            dist.broadcast(self.flat, src=src)
            for p, v in zip(self.params, self.views):
                p.copy_(v)

    def allreduce_(self) -> Optional[torch.Tensor]:
        """In place: p.grad <- mean over ranks of p.grad (a missing grad counts as zero)."""
        if not self.params:
            return None
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()
            else:
                v.copy_(p.grad)
        if self.world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.mul_(1.0 / self.world)
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                p.grad = v.clone()
            else:
                p.grad.copy_(v)
        return self.flat
