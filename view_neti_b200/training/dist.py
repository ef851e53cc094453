"""Batch-parallel plumbing: one process per GPU, ONE all-reduce per optimiser step over a flat fp32 buffer of the
mapper gradients (M_v + active M_o, ~1.1 MB) — SURVEY.md 8e.  Replaces the DDP wrapper that accelerate puts around the
text encoder (reference training/coach.py:97-99,214), which also fixes the reference's dict-held object mappers
escaping DDP (reference models/net_clip_text_embedding.py:25-32).  Works with the nccl (GPU) and gloo (CPU tests)
backends of torch.distributed.

Unused parameters keep the reference's semantics: a parameter whose .grad is None after backward on EVERY rank (the 13
inactive object mappers of a mode-3 step, coach.py:155-156 / dataset.py:584-600) keeps .grad None, so AdamW skips it
(no weight decay, no stale-momentum update) exactly as `optimizer.step()` does in the reference after `zero_grad()`.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist


class FlatGradAllReducer:
    """Packs the .grad of the given parameters into one persistent flat fp32 buffer, all-reduces it (mean) and
    scatters the result back.  One collective per call, whatever the number of tensors."""

    def __init__(self, params: Iterable[torch.nn.Parameter], device=None):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = device if device is not None else (self.params[0].device if self.params else "cpu")
        # gradients of all parameters + one "had a gradient" flag per parameter (summed by the same collective)
        self.flat = torch.zeros(n + len(self.params), dtype=torch.float32, device=dev)
        self.views, off = [], 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        self.n_grad = n
        self._index = {id(p): i for i, p in enumerate(self.params)}

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
            dist.broadcast(self.flat[:self.n_grad], src=src)
            for p, v in zip(self.params, self.views):
                p.copy_(v)

    def allreduce_(self, active: Optional[Sequence[torch.nn.Parameter]] = None) -> Optional[torch.Tensor]:
        """In place: p.grad <- mean over ranks of p.grad.

        active = None: every parameter takes part; one that has no gradient on this rank counts as zero, and one that has
        no gradient on ANY rank keeps .grad None (the flags travel in the tail of the same buffer; a rank reads them back
        - one small D2H copy - only if it holds such a parameter).
        active = [...]: the caller states which parameters the step used - the SAME list on every rank (mode 3: M_v + the
        object mapper that rank 0 drew for this step).  Only those are packed (a fixed-size prefix of the buffer), the
        others are left untouched: no flags, no read-back.
        With one process there is nothing to reduce and gradients are left exactly as backward produced them."""
        if not self.params or self.world == 1:
            return None
        world = self.world
        if active is not None:
            act = [p for p in active if id(p) in self._index]
            n = sum(p.numel() for p in act)
            buf = self.flat[:n]
            views, off = [], 0
            for p in act:
                views.append(buf[off:off + p.numel()].view_as(p))
                off += p.numel()
            for p, v in zip(act, views):
                if p.grad is None:
                    v.zero_()
                else:
                    v.copy_(p.grad)
            dist.all_reduce(buf, op=dist.ReduceOp.SUM)
            buf.mul_(1.0 / world)
            for p, v in zip(act, views):
                if p.grad is None:
                    p.grad = v.clone()
                else:
                    p.grad.copy_(v)
            return buf
        flags = self.flat[self.n_grad:]
        missing = [i for i, p in enumerate(self.params) if p.grad is None]
        flags.copy_(torch.tensor([0.0 if p.grad is None else 1.0 for p in self.params]), non_blocking=True)
        for i in missing:
            self.views[i].zero_()
        for p, v in zip(self.params, self.views):
            if p.grad is not None:
                v.copy_(p.grad)
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
        self.flat[:self.n_grad].mul_(1.0 / world)
        used_elsewhere = set()
        if missing:
            f = flags.tolist()
            used_elsewhere = {i for i in missing if f[i] > 0.0}
        for i, (p, v) in enumerate(zip(self.params, self.views)):
            if p.grad is not None:
                p.grad.copy_(v)
            elif i in used_elsewhere:
                p.grad = v.clone()
        return self.flat[:self.n_grad]
