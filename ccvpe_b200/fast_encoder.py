"""Inference-only execution plan for the (PyTorch) EfficientNet-B0 encoders of the bf16 throughput path.

The encoders stay PyTorch/cuDNN/cuBLAS (BASELINE.json north_star), but eager execution of the reference structure
spends most of its time in memory-bound glue (measured on B200, `profiles/r01_launches_bench_steps2.csv`: BatchNorm 35 %,
broadcast multiplies / padding copies 34 %, SiLU 8 %, all convolutions together < 12 %).  This module re-expresses the
same arithmetic with far fewer passes over the (6x expanded) activations -- algebra only, no new kernels:

  * BatchNorm (eval) is folded into the preceding convolution's weight and bias;
  * 1x1 convolutions on channels-last tensors are plain GEMMs (`F.linear`, bias in the cuBLAS epilogue);
  * the squeeze-excite gate is folded into the 1x1 projection:  W (g (.) x) = (W diag(g)) x  -> one `baddbmm` per block
    instead of a broadcast multiply over the expanded tensor followed by a convolution;
  * TensorFlow-style "same" padding (asymmetric for stride 2) and the ground encoder's circular width padding are
    produced by writing the preceding SiLU straight into the interior of a persistent pre-zeroed padded buffer
    (`aten::silu.out` on a strided view) instead of an `F.pad` copy;
  * symmetric zero padding of stride-1 depthwise convs is the convolution's own `padding` argument.

Numerics: same real-valued function as `efficientnet.EfficientNetB0` in eval mode; in fp32 it agrees to ~1e-5
(tests/test_fast_encoder.py), in bf16 it is the encoder of the bf16 path (tolerance stated in tests/test_gpu_forward.py).
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
from torch.nn import functional as F

from .efficientnet import EfficientNetB0, MBConv, SamePadConv2d


def _fold(conv: torch.nn.Conv2d, bn: torch.nn.BatchNorm2d) -> Tuple[torch.Tensor, torch.Tensor]:
    """conv (no bias) followed by eval-mode BN  ->  (weight', bias') in fp32."""
    w = conv.weight.detach().float()
    scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    bias = bn.bias.detach().float() - bn.running_mean.detach().float() * scale
    if conv.bias is not None:
        bias = bias + conv.bias.detach().float() * scale
    return w * scale.view(-1, 1, 1, 1), bias


class _Block:
    __slots__ = ("has_expand", "residual", "stride", "kernel", "pad_lo", "pad_hi", "w_exp", "b_exp", "w_dw", "b_dw",
                 "w_red", "b_red", "w_se", "b_se", "w_proj", "b_proj", "mid", "cout")


class FastEncoder:
    def __init__(self, enc: EfficientNetB0, dtype: torch.dtype = torch.bfloat16):
        self.dtype = dtype
        self.circular = bool(enc._conv_stem._circular)
        dev = enc._conv_stem.weight.device
        cast = lambda t: t.to(device=dev, dtype=dtype).contiguous()
        w, b = _fold(enc._conv_stem, enc._bn0)
        self.stem_w = cast(w).contiguous(memory_format=torch.channels_last)
        self.stem_b = cast(b)
        self.stem_pad = (enc._conv_stem._pad_lo, enc._conv_stem._pad_hi)
        self.blocks: List[_Block] = []
        for blk in enc._blocks:
            assert isinstance(blk, MBConv)
            o = _Block()
            o.has_expand, o.residual = blk._has_expand, blk._residual
            dw: SamePadConv2d = blk._depthwise_conv
            o.stride, o.kernel = dw.stride[0], dw.kernel_size[0]
            o.pad_lo, o.pad_hi = dw._pad_lo, dw._pad_hi
            if o.has_expand:
                w, b = _fold(blk._expand_conv, blk._bn0)
                o.w_exp, o.b_exp = cast(w.flatten(1)), cast(b)                       # [mid, cin]
            else:
                o.w_exp = o.b_exp = None
            w, b = _fold(dw, blk._bn1)
            o.w_dw, o.b_dw = cast(w).contiguous(memory_format=torch.channels_last), cast(b)
            o.mid = w.shape[0]
            o.w_red, o.b_red = cast(blk._se_reduce.weight.detach().flatten(1)), cast(blk._se_reduce.bias.detach())
            o.w_se, o.b_se = cast(blk._se_expand.weight.detach().flatten(1)), cast(blk._se_expand.bias.detach())
            w, b = _fold(blk._project_conv, blk._bn2)
            o.w_proj, o.b_proj = cast(w.flatten(1)), cast(b)                         # [cout, mid]
            o.cout = w.shape[0]
            self.blocks.append(o)
        w, b = _fold(enc._conv_head, enc._bn1)
        self.head_w, self.head_b = cast(w.flatten(1)), cast(b)
        self._buffers: Dict[tuple, torch.Tensor] = {}

    # -- padded staging buffers (borders zeroed once, interior rewritten on every use) ---------------------------
    def _padded(self, B, C, H, W, lo, hi, tag) -> torch.Tensor:
        key = (tag, B, C, H, W, lo, hi)
        buf = self._buffers.get(key)
        if buf is None:
            buf = torch.zeros((B, H + lo + hi, W + lo + hi, C), dtype=self.dtype, device=self.stem_w.device)
            self._buffers[key] = buf
        return buf                                                                   # NHWC physical

    def _silu_into_padded(self, y_nhwc: torch.Tensor, lo: int, hi: int, tag) -> torch.Tensor:
        """SiLU(y) written into the interior of a padded NHWC buffer; returns it as an NCHW-logical channels-last view."""
        B, H, W, C = y_nhwc.shape
        buf = self._padded(B, C, H, W, lo, hi, tag)
        torch.ops.aten.silu.out(y_nhwc, out=buf[:, lo:lo + H, lo:lo + W, :])
        if self.circular:                                  # wrap-around columns (vertical borders stay zero)
            if lo:
                buf[:, lo:lo + H, :lo, :] = buf[:, lo:lo + H, W:W + lo, :]
            if hi:
                buf[:, lo:lo + H, lo + W:, :] = buf[:, lo:lo + H, lo:lo + hi, :]
        return buf.permute(0, 3, 1, 2)

    # -- forward ------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def extract(self, x: torch.Tensor, keep_blocks: bool):
        """x: [B,3,H,W] (any format).  Returns (head features NCHW-logical/channels-last, [16 block outputs] or [])."""
        dt = self.dtype
        x = x.to(dt).contiguous(memory_format=torch.channels_last)
        lo, hi = self.stem_pad
        if self.circular:
            x = F.pad(F.pad(x, (lo, hi, 0, 0), mode="circular"), (0, 0, lo, hi))     # 3-channel input: negligible
        else:
            x = F.pad(x, (lo, hi, lo, hi))
        y = F.conv2d(x, self.stem_w, self.stem_b, stride=2)                          # [B,32,H/2,W/2] channels-last
        cur = y.permute(0, 2, 3, 1)                                                  # NHWC view, pre-activation
        pending_silu = True                                                          # `cur` still needs SiLU applied
        outs: List[torch.Tensor] = []
        for bi, o in enumerate(self.blocks):
            B, H, W, Cin = cur.shape
            if pending_silu and (o.has_expand or o.residual):
                cur = F.silu(cur)                                                    # needed as GEMM input / residual
                pending_silu = False
            block_in = cur
            if o.has_expand:
                e = F.linear(cur.reshape(B * H * W, Cin), o.w_exp, o.b_exp).view(B, H, W, o.mid)
                needs_silu = True
            else:
                e = cur
                needs_silu = pending_silu
            # depthwise conv input = SiLU(e) with TF-"same" padding
            use_buffer = self.circular or o.stride == 2
            if use_buffer:
                if needs_silu:
                    xin = self._silu_into_padded(e, o.pad_lo, o.pad_hi, "dw")
                else:
                    buf = self._padded(B, o.mid, H, W, o.pad_lo, o.pad_hi, "dw")
                    buf[:, o.pad_lo:o.pad_lo + H, o.pad_lo:o.pad_lo + W, :] = e
                    if self.circular:
                        if o.pad_lo:
                            buf[:, o.pad_lo:o.pad_lo + H, :o.pad_lo, :] = buf[:, o.pad_lo:o.pad_lo + H, W:W + o.pad_lo, :]
                        if o.pad_hi:
                            buf[:, o.pad_lo:o.pad_lo + H, o.pad_lo + W:, :] = \
                                buf[:, o.pad_lo:o.pad_lo + H, o.pad_lo:o.pad_lo + o.pad_hi, :]
                    xin = buf.permute(0, 3, 1, 2)
                d = F.conv2d(xin, o.w_dw, o.b_dw, stride=o.stride, groups=o.mid)
            else:
                xin = (F.silu(e) if needs_silu else e).permute(0, 3, 1, 2)
                d = F.conv2d(xin, o.w_dw, o.b_dw, stride=1, padding=o.pad_lo, groups=o.mid)
            pending_silu = False
            d = F.silu(d.permute(0, 2, 3, 1))                                        # [B,Ho,Wo,mid] NHWC
            Bo, Ho, Wo, _ = d.shape
            # squeeze-excite gate folded into the projection weights
            s = d.mean(dim=(1, 2), dtype=torch.float32).to(dt)                       # [B, mid]
            g = torch.sigmoid(F.linear(F.silu(F.linear(s, o.w_red, o.b_red)), o.w_se, o.b_se))   # [B, mid]
            wg = o.w_proj.unsqueeze(0) * g.unsqueeze(1)                              # [B, cout, mid]
            y = torch.baddbmm(o.b_proj.view(1, 1, -1), d.reshape(Bo, Ho * Wo, o.mid), wg.transpose(1, 2))
            y = y.view(Bo, Ho, Wo, o.cout)
            if o.residual:
                y = y + block_in
            cur = y
            if keep_blocks:
                outs.append(cur.permute(0, 3, 1, 2))
        B, H, W, C = cur.shape
        head = F.silu(F.linear(cur.reshape(B * H * W, C), self.head_w, self.head_b)).view(B, H, W, -1)
        return head.permute(0, 3, 1, 2), outs

    def extract_features(self, x):
        return self.extract(x, False)[0]

    def extract_features_multiscale(self, x):
        return self.extract(x, True)
