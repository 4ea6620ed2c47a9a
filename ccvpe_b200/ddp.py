"""Data-parallel training plumbing for BASELINE.json configs[4]: one process per GPU, weights replicated, the batch of
independent pairs split across ranks, and ONE collective -- the gradient all-reduce over NCCL / NVLink (north_star:
"NCCL over NVLink is used only for the DDP gradient allreduce in the training config").

`GradientAllReducer` is a small, explicit replacement of torch's DistributedDataParallel for this model:

  * every trainable parameter's `.grad` is a VIEW into one of a few flat fp32 buckets (filled in the order gradients
    become ready: the 98 head / decoder tensors arrive together when the CUDA post-encoder backward returns, then the
    aerial and ground encoders layer by layer from PyTorch autograd);
  * a post-accumulate-grad hook counts a bucket's parameters; when the last one is ready the bucket is all-reduced
    asynchronously (NCCL's own stream), overlapping the rest of the backward pass -- with the decoder's 43 M gradients
    in flight while the encoders' backward still runs;
  * `finish()` waits for the outstanding buckets (and averages); `zero_grad()` clears the buckets with one memset each.

The four `_fc.*` tensors of the encoders never receive a gradient (reference models.py:151,166 never call `_fc`), so they
are frozen here (DDP would otherwise need find_unused_parameters).  BatchNorm statistics stay per replica,
exactly like running the reference's single-GPU script on each shard (it has no SyncBN).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.distributed as dist


def freeze_unused(model: torch.nn.Module) -> List[str]:
    """requires_grad=False for parameters no forward of the reference ever touches (`*_efficientnet._fc.*`)."""
    frozen = []
    for name, p in model.named_parameters():
        if "._fc." in name:
            p.requires_grad_(False)
            frozen.append(name)
    return frozen


def ready_order(model: torch.nn.Module) -> List[torch.nn.Parameter]:
    """Trainable parameters in the order their gradients become available in a backward pass: decoder / head tensors
    first (they come out of one autograd Function), then each encoder from its last layer to its first."""
    named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
    dec = [p for n, p in named if not n.startswith(("grd_efficientnet.", "sat_efficientnet."))]
    sat = [p for n, p in named if n.startswith("sat_efficientnet.")]
    grd = [p for n, p in named if n.startswith("grd_efficientnet.")]
    return dec + sat[::-1] + grd[::-1]


class GradientAllReducer:
    def __init__(self, model: torch.nn.Module, process_group=None, bucket_mb: float = 48.0, first_bucket_mb: float = 48.0):
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.frozen = freeze_unused(model)
        params = ready_order(model)
        if not params:
            raise ValueError("no trainable parameters")
        self.device = params[0].device
        self.buckets: List[torch.Tensor] = []
        self._bucket_of: Dict[int, int] = {}
        self._members: List[List[torch.nn.Parameter]] = []
        cap = int(first_bucket_mb * (1 << 20)) // 4
        cur: List[torch.nn.Parameter] = []
        cur_n = 0
        for p in params:
            if cur and cur_n + p.numel() > cap:
                self._seal(cur, cur_n)
                cur, cur_n = [], 0
                cap = int(bucket_mb * (1 << 20)) // 4
            cur.append(p)
            cur_n += p.numel()
        if cur:
            self._seal(cur, cur_n)
        self._pending = [0] * len(self.buckets)
        self._works: List[Optional[object]] = [None] * len(self.buckets)
        self._launched = [False] * len(self.buckets)
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in params]
        self.n_params = sum(p.numel() for p in params)
        self._avg_native = self.world > 1 and dist.get_backend(process_group) == "nccl"
        self.zero_grad()

    def _seal(self, members, n):
        flat = torch.zeros(n, dtype=torch.float32, device=self.device)
        off = 0
        for p in members:
            if p.dtype != torch.float32:
                raise TypeError("GradientAllReducer expects fp32 master parameters")
            p.grad = flat[off:off + p.numel()].view_as(p)
            self._bucket_of[id(p)] = len(self.buckets)
            off += p.numel()
        self.buckets.append(flat)
        self._members.append(list(members))

    # -- per step -----------------------------------------------------------------------------------------------
    def zero_grad(self):
        """Clears every bucket (the parameters' .grad are views into them) and re-arms the hooks."""
        for b, flat in enumerate(self.buckets):
            flat.zero_()
            self._pending[b] = len(self._members[b])
            self._works[b] = None
            self._launched[b] = False

    def _on_grad(self, p: torch.nn.Parameter):
        b = self._bucket_of[id(p)]
        if p.grad.data_ptr() < self.buckets[b].data_ptr() or p.grad.data_ptr() >= self.buckets[b].data_ptr() + \
                self.buckets[b].numel() * 4:
            # someone replaced .grad (e.g. optimizer.zero_grad(set_to_none=True)): fold it back into the bucket view
            off = sum(q.numel() for q in self._members[b][: [id(q) for q in self._members[b]].index(id(p))])
            view = self.buckets[b][off:off + p.numel()].view_as(p)
            view.copy_(p.grad)
            p.grad = view
        self._pending[b] -= 1
        if self._pending[b] == 0:
            self._launch(b)

    def _launch(self, b: int):
        self._launched[b] = True
        if self.world == 1:
            return
        op = dist.ReduceOp.AVG if self._avg_native else dist.ReduceOp.SUM
        self._works[b] = dist.all_reduce(self.buckets[b], op=op, group=self.group, async_op=True)

    def finish(self):
        """Blocks the current stream until every bucket is reduced (buckets whose hooks never all fired -- parameters
        without a gradient this step -- are reduced now) and averages over the ranks."""
        for b in range(len(self.buckets)):
            if not self._launched[b]:
                self._launch(b)
        for b, w in enumerate(self._works):
            if w is not None:
                w.wait()
                if not self._avg_native:
                    self.buckets[b].div_(self.world)
                self._works[b] = None

    def all_reduce_alone(self):
        """Diagnostic: all-reduce every bucket back to back with nothing overlapped (bench: exposed vs alone)."""
        if self.world == 1:
            return
        works = [dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True) for flat in self.buckets]
        for w in works:
            w.wait()

    @property
    def bucket_bytes(self) -> List[int]:
        return [int(b.numel()) * 4 for b in self.buckets]

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []
