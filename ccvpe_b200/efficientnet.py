"""EfficientNet-B0 ground / aerial encoder (stays in PyTorch per BASELINE.json north_star).

This is the one piece of the model that is NOT rebuilt as CUDA kernels: it is plain PyTorch so that
cuDNN runs it.  It is written from the published EfficientNet-B0 architecture table and made
state_dict-compatible with the encoder the reference vendors, so that a reference checkpoint loads
strictly:

  * parameter / buffer names follow reference `efficientnet_pytorch/model.py:48-88, 162-219`
    (`_conv_stem`, `_bn0`, `_blocks.N.{_expand_conv,_bn0,_depthwise_conv,_bn1,_se_reduce,_se_expand,
    _project_conv,_bn2}`, `_conv_head`, `_bn1`, `_fc`);
  * "same" padding is the *static* TensorFlow-style padding the reference computes for a 224x224
    image regardless of the real input size (`utils.py:254-282`, `model.py:175-176`): stride-1 convs pad
    (k-1)/2 on both sides, stride-2 convs pad (k-2)//2 before and k-2-(k-2)//2 after;
  * `circular=True` replaces the horizontal zero padding by wrap-around padding (panorama azimuth is
    periodic), vertical padding stays zero (`utils.py:330-358`);
  * activation is x*sigmoid(x) evaluated exactly like the reference (`utils.py:64-68`) unless
    `fast_activation` is set, in which case the fused `F.silu` kernel is used (differs in the last ulp);
  * stochastic depth draws `torch.rand([B,1,1,1])` per residual block in training mode in the same
    order as the reference (`utils.py:129-154`, `model.py:124-130`) so RNG streams stay aligned.
"""
from __future__ import annotations

import math
from typing import List, Tuple

import torch
from torch import nn
from torch.nn import functional as F

# (repeats, kernel, stride, expand, in, out) -- EfficientNet-B0, Tan & Le 2019, table 1.
_B0_STAGES: Tuple[Tuple[int, int, int, int, int, int], ...] = (
    (1, 3, 1, 1, 32, 16),
    (2, 3, 2, 6, 16, 24),
    (2, 5, 2, 6, 24, 40),
    (3, 3, 2, 6, 40, 80),
    (3, 5, 1, 6, 80, 112),
    (4, 5, 2, 6, 112, 192),
    (1, 3, 1, 6, 192, 320),
)
_SE_RATIO = 0.25
_BN_MOMENTUM = 0.01  # 1 - 0.99 (TF momentum convention)
_BN_EPS = 1e-3
_DROP_CONNECT = 0.2
_HEAD_CHANNELS = 1280
_STATIC_IMAGE = 224

#: indices of the block outputs the decoders use as skips (reference models.py:167-171)
SKIP_BLOCKS = (0, 2, 4, 10, 15)


class _Swish(nn.Module):
    def __init__(self):
        super().__init__()
        self.fast = False

    def forward(self, x):
        if self.fast:
            return F.silu(x)
        return x * torch.sigmoid(x)


class SamePadConv2d(nn.Conv2d):
    """Conv2d with the reference's static 'same' padding; optionally circular along width."""

    def __init__(self, cin, cout, kernel, stride=1, groups=1, bias=False, circular=False, static_size=_STATIC_IMAGE):
        super().__init__(cin, cout, kernel, stride=stride, padding=0, groups=groups, bias=bias)
        out = math.ceil(static_size / stride)
        total = max((out - 1) * stride + kernel - static_size, 0)
        self._pad_lo = total // 2
        self._pad_hi = total - total // 2
        self._circular = bool(circular)

    def forward(self, x):
        lo, hi = self._pad_lo, self._pad_hi
        if lo or hi:
            # (pure data movement, so any formulation is bit-identical to the reference's F.pad calls; this one keeps a
            # channels-last input channels-last -- F.pad(mode="circular") would hand back an NCHW tensor and push the
            # convolution and the BatchNorm behind it onto the slow NCHW kernels in training)
            cl = x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last) and not x.is_contiguous()
            if self._circular:
                w = x.shape[3]
                parts = ([x[..., w - lo:]] if lo else []) + [x] + ([x[..., :hi]] if hi else [])
                x = torch.cat(parts, dim=3) if len(parts) > 1 else x
                x = F.pad(x, (0, 0, lo, hi))
            else:
                x = F.pad(x, (lo, hi, lo, hi))
            if cl and not x.is_contiguous(memory_format=torch.channels_last):
                x = x.contiguous(memory_format=torch.channels_last)
        return F.conv2d(x, self.weight, self.bias, self.stride, 0, self.dilation, self.groups)


class MBConv(nn.Module):
    def __init__(self, kernel, stride, expand, cin, cout, circular, static_size):
        super().__init__()
        mid = cin * expand
        self._has_expand = expand != 1
        self._residual = stride == 1 and cin == cout
        if self._has_expand:
            self._expand_conv = SamePadConv2d(cin, mid, 1, circular=circular, static_size=static_size)
            self._bn0 = nn.BatchNorm2d(mid, momentum=_BN_MOMENTUM, eps=_BN_EPS)
        self._depthwise_conv = SamePadConv2d(mid, mid, kernel, stride=stride, groups=mid, circular=circular,
                                             static_size=static_size)
        self._bn1 = nn.BatchNorm2d(mid, momentum=_BN_MOMENTUM, eps=_BN_EPS)
        squeezed = max(1, int(cin * _SE_RATIO))
        self._se_reduce = nn.Conv2d(mid, squeezed, 1)
        self._se_expand = nn.Conv2d(squeezed, mid, 1)
        self._project_conv = SamePadConv2d(mid, cout, 1, circular=circular,
                                           static_size=math.ceil(static_size / stride))
        self._bn2 = nn.BatchNorm2d(cout, momentum=_BN_MOMENTUM, eps=_BN_EPS)
        self._swish = _Swish()

    def forward(self, x, drop_rate: float):
        y = x
        if self._has_expand:
            y = self._swish(self._bn0(self._expand_conv(y)))
        y = self._swish(self._bn1(self._depthwise_conv(y)))
        gate = F.adaptive_avg_pool2d(y, 1)
        gate = self._se_expand(self._swish(self._se_reduce(gate)))
        y = torch.sigmoid(gate) * y
        y = self._bn2(self._project_conv(y))
        if self._residual:
            if drop_rate and self.training:
                keep = 1.0 - drop_rate
                mask = torch.floor(keep + torch.rand([y.shape[0], 1, 1, 1], dtype=y.dtype, device=y.device))
                y = y / keep * mask
            y = y + x
        return y


class EfficientNetB0(nn.Module):
    """EfficientNet-B0 trunk exposing the two feature extractors the CVM_* models call."""

    def __init__(self, circular: bool = False, num_classes: int = 1000):
        super().__init__()
        size = _STATIC_IMAGE
        self._conv_stem = SamePadConv2d(3, 32, 3, stride=2, circular=circular, static_size=size)
        self._bn0 = nn.BatchNorm2d(32, momentum=_BN_MOMENTUM, eps=_BN_EPS)
        size = math.ceil(size / 2)
        blocks: List[MBConv] = []
        for repeats, kernel, stride, expand, cin, cout in _B0_STAGES:
            for r in range(repeats):
                blocks.append(MBConv(kernel, stride if r == 0 else 1, expand, cin if r == 0 else cout, cout,
                                     circular, size))
                if r == 0:
                    size = math.ceil(size / stride)
        self._blocks = nn.ModuleList(blocks)
        self._conv_head = SamePadConv2d(_B0_STAGES[-1][5], _HEAD_CHANNELS, 1, circular=circular, static_size=size)
        self._bn1 = nn.BatchNorm2d(_HEAD_CHANNELS, momentum=_BN_MOMENTUM, eps=_BN_EPS)
        # present in the reference state_dict (include_top=True) although no CVM_* forward ever uses it
        self._fc = nn.Linear(_HEAD_CHANNELS, num_classes)
        self._swish = _Swish()

    # -- configuration -------------------------------------------------------------------------
    def set_fast_activation(self, fast: bool = True):
        for m in self.modules():
            if isinstance(m, _Swish):
                m.fast = bool(fast)
        return self

    # -- the two entry points of reference model.py:278-326 -------------------------------------
    def _trunk(self, x, keep_blocks: bool):
        x = self._swish(self._bn0(self._conv_stem(x)))
        n = len(self._blocks)
        outs = []
        for i, block in enumerate(self._blocks):
            x = block(x, _DROP_CONNECT * float(i) / n)
            if keep_blocks:
                outs.append(x)
        x = self._swish(self._bn1(self._conv_head(x)))
        return x, outs

    def extract_features(self, inputs):
        return self._trunk(inputs, False)[0]

    def extract_features_multiscale(self, inputs):
        return self._trunk(inputs, True)

    def forward(self, inputs):
        x = self.extract_features(inputs)
        x = F.adaptive_avg_pool2d(x, 1).flatten(1)
        return self._fc(F.dropout(x, 0.2, self.training))
