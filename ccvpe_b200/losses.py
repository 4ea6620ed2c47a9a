"""Drop-in replacements of the reference's training losses (reference losses.py:4-29), each ONE fused CUDA call that
produces the loss value and its gradient (libccvpe_b200: ccvpe_infonce_loss / ccvpe_cross_entropy_loss /
ccvpe_orientation_loss).  Same names, argument order and semantics as the reference, so train_VIGOR.py:137-146 reads

    loss_ori = orientation_loss(ori, gt_orientation, gt)
    loss_infoNCE = infoNCELoss(torch.flatten(matching_score_stacked, start_dim=1), torch.flatten(gt_bottleneck, start_dim=1))
    loss_ce = cross_entropy_loss(logits_flattened, gt_flattened)

unchanged.  Labels / ground-truth maps are treated as constants (they are data in the reference too); there is no CPU path.
"""
from __future__ import annotations

import torch

from . import cabi


def _f32c(t: torch.Tensor) -> torch.Tensor:
    return t.detach().contiguous().float()


class _InfoNCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, scores, labels, temperature):
        if not scores.is_cuda:
            raise cabi.CcvpeError("ccvpe_b200.losses run on CUDA only; there is no CPU fallback")
        s, l = _f32c(scores), _f32c(labels)
        loss = torch.empty((), dtype=torch.float32, device=s.device)
        ds = torch.empty_like(s)
        with cabi.device_of(s):
            cabi.infonce_loss(s, l, float(temperature), loss, ds)
        ctx.save_for_backward(ds)
        return loss

    @staticmethod
    def backward(ctx, g):
        (ds,) = ctx.saved_tensors
        return ds * g, None, None


class _CrossEntropy(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels):
        if not logits.is_cuda:
            raise cabi.CcvpeError("ccvpe_b200.losses run on CUDA only; there is no CPU fallback")
        x, l = _f32c(logits), _f32c(labels)
        loss = torch.empty((), dtype=torch.float32, device=x.device)
        dx = torch.empty_like(x)
        with cabi.device_of(x):
            cabi.cross_entropy_loss(x, l, loss, dx)
        ctx.save_for_backward(dx)
        return loss

    @staticmethod
    def backward(ctx, g):
        (dx,) = ctx.saved_tensors
        return dx * g, None


class _Orientation(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ori, gt_orientation, gt):
        if not ori.is_cuda:
            raise cabi.CcvpeError("ccvpe_b200.losses run on CUDA only; there is no CPU fallback")
        o, go, w = _f32c(ori), _f32c(gt_orientation), _f32c(gt)
        loss = torch.empty((), dtype=torch.float32, device=o.device)
        do = torch.empty_like(o)
        with cabi.device_of(o):
            cabi.orientation_loss(o, go, w, loss, do)
        ctx.save_for_backward(do)
        return loss

    @staticmethod
    def backward(ctx, g):
        (do,) = ctx.saved_tensors
        return do * g, None, None


def infoNCELoss(scores, labels, temperature=0.1):
    """reference losses.py:4-20 -- weighted InfoNCE over a flattened score volume [B, n]; positives = labels > 1e-2."""
    return _InfoNCE.apply(scores, labels, temperature)


def cross_entropy_loss(logits, labels):
    """reference losses.py:23-24 -- -sum(labels * log_softmax(logits, dim=1)) / B."""
    return _CrossEntropy.apply(logits, labels)


def orientation_loss(ori, gt_orientation, gt):
    """reference losses.py:28-29 -- sum(sum((gt_orientation - ori)^2, dim=1, keepdim=True) * gt) / B;
    ori, gt_orientation [B, 2, H, W], gt [B, 1, H, W]."""
    return _Orientation.apply(ori, gt_orientation, gt)


def training_loss(outputs, gt, gt_with_ori, gt_orientation, weight_infoNCE=1e4, weight_ori=1e1):
    """The reference's loss combination (train_VIGOR.py:120-146) on the 9-tuple a model returns: GT preparation
    (flatten + normalise `gt`; max-pool `gt_with_ori` to the six score-volume resolutions) and
    loss_ce + weight_infoNCE * mean(infoNCE_1..6) + weight_ori * loss_ori."""
    logits_flattened, _heatmap, ori = outputs[0], outputs[1], outputs[2]
    gt_flattened = torch.flatten(gt, start_dim=1)
    gt_flattened = gt_flattened / torch.sum(gt_flattened, dim=1, keepdim=True)
    loss_ori = orientation_loss(ori, gt_orientation, gt)
    nce = 0
    for scores, k in zip(outputs[3:9], (64, 32, 16, 8, 4, 2)):
        gt_b = torch.nn.functional.max_pool2d(gt_with_ori, k, stride=k)
        nce = nce + infoNCELoss(torch.flatten(scores, start_dim=1), torch.flatten(gt_b, start_dim=1))
    loss_ce = cross_entropy_loss(logits_flattened, gt_flattened)
    return loss_ce + weight_infoNCE * nce / 6 + weight_ori * loss_ori
