"""Inference-only execution plan for the (PyTorch) EfficientNet-B0 encoders of the bf16 throughput path.

The encoders stay PyTorch/cuDNN/cuBLAS (BASELINE.json north_star), but eager execution of the reference structure
spends most of its time in memory-bound glue (measured on B200, `profiles/r01_launches_bench_steps2.csv`: BatchNorm 35 %,
broadcast multiplies / padding copies 34 %, SiLU 8 %, all convolutions together < 12 %).  This module re-expresses the
same arithmetic with far fewer passes over the (6x expanded) activations -- algebra only, no new kernels:

  * BatchNorm (eval) is folded into the preceding convolution's weight and bias;
  * 1x1 convolutions on channels-last tensors are plain GEMMs; on the CUDA bf16 path the expand / head GEMMs run on
    libccvpe_b200's tcgen05 pipeline with bias + SiLU in the epilogue, stored straight into the depthwise conv's padded
    input image (`ccvpe_pointwise_silu_nhwc`), elsewhere `F.linear`;
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

from . import cabi
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
    __slots__ = ("has_expand", "residual", "stride", "kernel", "pad_lo", "pad_hi", "w_exp", "b_exp", "w_exp_nk",
                 "b_exp_f32", "w_dw", "b_dw",
                 "w_red", "b_red", "w_se", "w_se_t", "b_se", "w_proj", "b_proj", "mid", "cout", "w_dw_taps")


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
        # fp32 operands of libccvpe_b200's fused stem kernel: weights [27, 32] ordered (ci, ky, kx)
        self.stem_w_taps = w.permute(1, 2, 3, 0).reshape(27, -1).to(device=dev, dtype=torch.float32).contiguous()
        self.stem_b_f32 = b.to(device=dev, dtype=torch.float32).contiguous()
        self.blocks: List[_Block] = []
        # The projection bias (folded BN shift) of every block is a constant per-channel vector.  Instead of adding it
        # (a broadcast pass per block) it is carried as a "pending" constant: folded into the next expand GEMM's bias
        # (W_e (y + p) + b_e = W_e y + (W_e p + b_e)), accumulated through residual connections (which become the GEMM's
        # beta = 1 accumulate), and only materialised where a tensor leaves the encoder (decoder skips, head).
        pending = None                                                               # fp32 [C] or None
        for blk in enc._blocks:
            assert isinstance(blk, MBConv)
            o = _Block()
            o.has_expand, o.residual = blk._has_expand, blk._residual
            dw: SamePadConv2d = blk._depthwise_conv
            o.stride, o.kernel = dw.stride[0], dw.kernel_size[0]
            o.pad_lo, o.pad_hi = dw._pad_lo, dw._pad_hi
            if o.has_expand:
                w, b = _fold(blk._expand_conv, blk._bn0)
                if pending is not None:
                    b = b + w.flatten(1) @ pending.to(w.device)
                o.w_exp, o.b_exp = cast(w.flatten(1)), cast(b)                       # [mid, cin]
                # operands of libccvpe_b200's tcgen05 pointwise conv (+ bias + SiLU epilogue)
                o.w_exp_nk = cabi.pad_k_blocks(o.w_exp)
                o.b_exp_f32 = b.to(device=dev, dtype=torch.float32).contiguous()
            else:
                assert pending is None                                               # only the first block has no expand
                o.w_exp = o.b_exp = None
            w, b = _fold(dw, blk._bn1)
            o.w_dw, o.b_dw = cast(w).contiguous(memory_format=torch.channels_last), cast(b)
            o.mid = w.shape[0]
            o.w_dw_taps = cast(w.reshape(o.mid, -1).t())                             # [k*k, mid] for the fused kernel
            o.w_red, o.b_red = cast(blk._se_reduce.weight.detach().flatten(1)), cast(blk._se_reduce.bias.detach())
            o.w_se, o.b_se = cast(blk._se_expand.weight.detach().flatten(1)), cast(blk._se_expand.bias.detach())
            o.w_se_t = o.w_se.t().contiguous()                                       # [R, mid] for the fused gate kernel
            w, b = _fold(blk._project_conv, blk._bn2)
            pending = b + (pending if (o.residual and pending is not None) else 0.0)
            o.w_proj, o.b_proj = cast(w.flatten(1)), cast(pending)                   # [cout, mid]; b_proj = pending AFTER this block
            o.cout = w.shape[0]
            self.blocks.append(o)
        w, b = _fold(enc._conv_head, enc._bn1)
        b = b + w.flatten(1) @ pending.to(w.device)
        self.head_w, self.head_b = cast(w.flatten(1)), cast(b)
        self.head_w_nk = cabi.pad_k_blocks(self.head_w)
        self.head_b_f32 = b.to(device=dev, dtype=torch.float32).contiguous()
        self._buffers: Dict[tuple, torch.Tensor] = {}
        self._mid_total = sum(o.mid for o in self.blocks)
        self.keep = set(range(len(self.blocks)))        # block outputs materialised by extract_features_multiscale

    # -- padded staging buffers (borders zeroed once, interior rewritten on every use) ---------------------------
    def _padded(self, B, C, H, W, lo, hi, tag) -> torch.Tensor:
        key = (tag, B, C, H, W, lo, hi)
        buf = self._buffers.get(key)
        if buf is None:
            buf = torch.zeros((B, H + lo + hi, W + lo + hi, C), dtype=self.dtype, device=self.stem_w.device)
            self._buffers[key] = buf
        return buf                                                                   # NHWC physical

    def _packed_bias(self, o, rep) -> torch.Tensor:
        key = ("bproj", id(o), rep)
        buf = self._buffers.get(key)
        if buf is None:
            buf = o.b_proj.repeat(rep).contiguous()
            self._buffers[key] = buf
        return buf

    def _wrap_columns(self, buf, H, W, lo, hi):
        """Circular width padding: copy the wrap-around columns inside the padded buffer (vertical borders stay zero)."""
        if buf.is_cuda and buf.dtype == torch.bfloat16 and buf.is_contiguous():
            cabi.wrap_columns_nhwc(buf, H, W, lo, hi)
            return
        if lo:
            buf[:, lo:lo + H, :lo, :] = buf[:, lo:lo + H, W:W + lo, :]
        if hi:
            buf[:, lo:lo + H, lo + W:, :] = buf[:, lo:lo + H, lo:lo + hi, :]

    def _act(self, x_nhwc: torch.Tensor, bias, pad=None, want_sum: bool = False):
        """y = SiLU(x + bias) in ONE pass; optionally written into the interior of a padded staging buffer (pad = (lo, hi))
        and/or accumulating the per-(b, c) sums the squeeze-excite gate needs.  CUDA bf16: libccvpe_b200's fused kernel;
        otherwise (CPU / fp32 checks) the same arithmetic with torch ops.
        Returns (y as NHWC view, padded NCHW-logical view or None, channel sums fp32 [B, C] or None)."""
        B, H, W, C = x_nhwc.shape
        buf = None
        if pad is not None:
            lo, hi = pad
            buf = self._padded(B, C, H, W, lo, hi, "dw")
            out = buf[:, lo:lo + H, lo:lo + W, :]
        else:
            out = torch.empty((B, H, W, C), dtype=x_nhwc.dtype, device=x_nhwc.device)
        sums = None
        if x_nhwc.is_cuda and x_nhwc.dtype == torch.bfloat16:
            if want_sum:
                sums = torch.zeros((B, C), dtype=torch.int64, device=x_nhwc.device)
            cabi.bias_silu_nhwc(x_nhwc.contiguous(), bias, out, sums)
            if want_sum:
                sums = sums.double().div_(cabi.SE_SUM_SCALE).float()
        else:
            t = x_nhwc if bias is None else x_nhwc + bias
            torch.ops.aten.silu.out(t, out=out)
            if want_sum:
                sums = out.float().sum(dim=(1, 2))
        padded_view = None
        if buf is not None:
            if self.circular:
                self._wrap_columns(buf, H, W, pad[0], pad[1])
            padded_view = buf.permute(0, 3, 1, 2)
        return out, padded_view, sums

    # -- forward ------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def extract(self, x: torch.Tensor, keep_blocks: bool):
        """x: [B,3,H,W] (any format).  Returns (head features NCHW-logical/channels-last, [16 block outputs] or [])."""
        dt = self.dtype
        lo, hi = self.stem_pad
        fused_dw = x.is_cuda and dt == torch.bfloat16      # libccvpe_b200's kernels (stem, expand, depthwise) vs torch ops
        stem_padded = None
        if fused_dw and self.stem_w_taps.shape[1] == 32 and x.shape[1] == 3 and not self.blocks[0].has_expand:
            # stem conv + bias + SiLU straight from the fp32 NCHW image into block 0's padded depthwise input
            o0 = self.blocks[0]
            B, _c, H, W = x.shape
            Ho, Wo = (H + lo + hi - 3) // 2 + 1, (W + lo + hi - 3) // 2 + 1
            buf = self._padded(B, 32, Ho, Wo, o0.pad_lo, o0.pad_hi, "dw")
            if x.dtype == torch.uint8:
                # f4: ToTensor + ImageNet Normalize fused into the stem's loads (reference train_VIGOR.py:55-70)
                cabi.stem_conv_silu_u8_nhwc(x.contiguous(), self.stem_w_taps, self.stem_b_f32, buf, lo, hi, o0.pad_lo,
                                            o0.pad_hi, self.circular)
            else:
                cabi.stem_conv_silu_nhwc(x.float().contiguous(), self.stem_w_taps, self.stem_b_f32, buf, lo, hi, o0.pad_lo,
                                         o0.pad_hi, self.circular)
            stem_padded = buf.permute(0, 3, 1, 2)
            pre = pre_bias = None
        else:
            if x.dtype == torch.uint8:
                raise cabi.CcvpeError("uint8 images are ingested by the CUDA bf16 encoder plan only (models.ingest otherwise)")
            x = x.to(dt).contiguous(memory_format=torch.channels_last)
            if self.circular:
                x = F.pad(F.pad(x, (lo, hi, 0, 0), mode="circular"), (0, 0, lo, hi))     # 3-channel input: negligible
            else:
                x = F.pad(x, (lo, hi, lo, hi))
            pre = F.conv2d(x, self.stem_w, None, stride=2).permute(0, 2, 3, 1)          # NHWC, before bias + SiLU
            pre_bias = self.stem_b
        cur = None                                                                   # activated NHWC tensor
        outs: List[torch.Tensor] = []
        # squeeze-excite channel sums of all blocks: one zero fill per forward instead of one per block
        sums_ws = torch.zeros(x.shape[0] * self._mid_total, dtype=torch.int64, device=x.device) if fused_dw else None
        sums_off = 0
        for o in self.blocks:
            dw_pad = (o.pad_lo, o.pad_hi) if (self.circular or o.stride == 2 or fused_dw) else None
            if stem_padded is not None:
                block_in = None
                mid_in_plain, mid_in_padded = None, stem_padded
                stem_padded = None
            elif pre is not None and not o.has_expand and not o.residual:
                # block 0: the stem's pending bias + SiLU is applied straight into the depthwise conv's input
                plain, padded, _unused = self._act(pre, pre_bias, pad=dw_pad)
                block_in = None
                mid_in_plain, mid_in_padded = (plain if padded is None else None), padded
                pre = None
            else:
                if pre is not None:
                    cur = self._act(pre, pre_bias)[0]
                    pre = None
                block_in = cur
                B, H, W, Cin = cur.shape
                if o.has_expand and fused_dw:
                    # expand GEMM + bias + SiLU in one tcgen05 kernel, written straight into the depthwise conv's
                    # padded input image
                    buf = self._padded(B, o.mid, H, W, o.pad_lo, o.pad_hi, "dw")
                    cabi.pointwise_silu_nhwc(cur, o.w_exp_nk, o.b_exp_f32, buf, o.pad_lo, o.pad_hi)
                    if self.circular:
                        self._wrap_columns(buf, H, W, o.pad_lo, o.pad_hi)
                    mid_in_plain, mid_in_padded = None, buf.permute(0, 3, 1, 2)
                elif o.has_expand:
                    e = F.linear(cur.reshape(B * H * W, Cin), o.w_exp, o.b_exp).view(B, H, W, o.mid)
                    plain, padded, _unused = self._act(e, None, pad=dw_pad)
                    mid_in_plain, mid_in_padded = (plain if padded is None else None), padded
                else:
                    if dw_pad is not None:
                        buf = self._padded(B, o.mid, H, W, o.pad_lo, o.pad_hi, "dw")
                        buf[:, o.pad_lo:o.pad_lo + H, o.pad_lo:o.pad_lo + W, :] = cur
                        if self.circular:
                            self._wrap_columns(buf, H, W, o.pad_lo, o.pad_hi)
                        mid_in_plain, mid_in_padded = None, buf.permute(0, 3, 1, 2)
                    else:
                        mid_in_plain, mid_in_padded = cur, None
            if fused_dw:
                xp = mid_in_padded.permute(0, 2, 3, 1)                               # padded NHWC view
                Bp, Hp, Wp, _ = xp.shape
                Ho, Wo = (Hp - o.kernel) // o.stride + 1, (Wp - o.kernel) // o.stride + 1
                d = torch.empty((Bp, Ho, Wo, o.mid), dtype=dt, device=xp.device)
                sums = sums_ws[sums_off:sums_off + Bp * o.mid].view(Bp, o.mid)
                sums_off += Bp * o.mid
                cabi.dwconv_bias_silu_nhwc(xp, o.w_dw_taps, o.b_dw, d, o.kernel, o.stride, sums)
            else:
                if mid_in_padded is not None:
                    d = F.conv2d(mid_in_padded, o.w_dw, None, stride=o.stride, groups=o.mid)
                else:
                    d = F.conv2d(mid_in_plain.permute(0, 3, 1, 2), o.w_dw, None, stride=1, padding=o.pad_lo,
                                 groups=o.mid)
                d, _unused, sums = self._act(d.permute(0, 2, 3, 1), o.b_dw, want_sum=True)  # [B,Ho,Wo,mid] + SE squeeze
            Bo, Ho, Wo, _ = d.shape
            # squeeze-excite gate folded into the projection weights
            if fused_dw:
                # shallow first block (mid = 32: 64-byte pixels): two pixels per GEMM row with block-diagonal weights, so
                # the projection's TMA boxes carry full 128-byte rows (its load rate is per box row, not per byte)
                rep = 2 if (o.mid <= 32 and (Ho * Wo) % 2 == 0 and not o.residual) else 1
                wg = torch.empty((Bo, rep * o.cout, rep * o.mid), dtype=dt, device=d.device)
                gate_ws = torch.empty((Bo, o.mid), dtype=torch.float32, device=d.device)
                cabi.se_gate_scale(sums, 1.0 / float(Ho * Wo), o.w_red, o.b_red, o.w_se_t, o.b_se, o.w_proj, wg, gate_ws, rep)
            else:
                sq = (sums / float(Ho * Wo)).to(dt)                                  # [B, mid]
                g = torch.sigmoid(F.linear(F.silu(F.linear(sq, o.w_red, o.b_red)), o.w_se, o.b_se))   # [B, mid]
                wg = o.w_proj.unsqueeze(0) * g.unsqueeze(1)                          # [B, cout, mid]
            if fused_dw:
                # batched per-image-weight GEMM on tcgen05: the identity skip is added in the epilogue, and the copy the
                # decoder reads (cur + pending bias) is a second output of the same pass
                cur = torch.empty((Bo, Ho, Wo, o.cout), dtype=dt, device=d.device)
                want = keep_blocks and len(outs) in self.keep
                kept = torch.empty_like(cur) if want else None
                if rep > 1:
                    cabi.mbconv_project_nhwc(d.view(Bo, Ho * Wo // rep, rep * o.mid), wg, None,
                                             cur.view(Bo, Ho * Wo // rep, rep * o.cout),
                                             self._packed_bias(o, rep) if want else None,
                                             kept.view(Bo, Ho * Wo // rep, rep * o.cout) if want else None)
                else:
                    cabi.mbconv_project_nhwc(d, wg, block_in if o.residual else None, cur, o.b_proj if want else None, kept)
                if keep_blocks:
                    outs.append(kept.permute(0, 3, 1, 2) if want else None)
                continue
            dm = d.reshape(Bo, Ho * Wo, o.mid)
            if o.residual:                                                           # residual add = beta 1 of the GEMM
                y = torch.baddbmm(block_in.reshape(Bo, Ho * Wo, o.cout), dm, wg.transpose(1, 2))
            else:
                y = torch.bmm(dm, wg.transpose(1, 2))
            cur = y.view(Bo, Ho, Wo, o.cout)                                         # true value = cur + o.b_proj (pending)
            if keep_blocks:
                bi = len(outs)
                outs.append((cur + o.b_proj).permute(0, 3, 1, 2) if bi in self.keep else None)
        B, H, W, C = cur.shape
        if fused_dw:
            head = torch.empty((B, H, W, self.head_w.shape[0]), dtype=dt, device=cur.device)
            cabi.pointwise_silu_nhwc(cur, self.head_w_nk, self.head_b_f32, head)
        else:
            head_pre = F.linear(cur.reshape(B * H * W, C), self.head_w, self.head_b).view(B, H, W, -1)
            head = self._act(head_pre, None)[0]
        return head.permute(0, 3, 1, 2), outs

    def extract_features(self, x):
        return self.extract(x, False)[0]

    def extract_features_multiscale(self, x):
        return self.extract(x, True)
