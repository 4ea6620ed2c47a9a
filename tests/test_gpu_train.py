"""GPU parity of the training step (BASELINE.json configs[4]): every backward operator of the C ABI against torch autograd
of the CPU oracle on the same seeded inputs, then loss + every gradient of a whole CVM_VIGOR training step (reference
train_VIGOR.py:120-150) against autograd through the oracle.

Tolerances (stated per test): fp32 operators 2e-4 of max|ref| per tensor; whole step fp32: loss 1e-4 relative, gradients
2e-3 of max|ref| per tensor; bf16 step: cosine similarity of every large gradient tensor with the fp32 oracle's >= 0.98."""
import copy
import math

import pytest
import torch
from torch.nn import functional as F

from ccvpe_b200 import cabi, losses
from ccvpe_b200.synthetic import synthetic_ground_truth
from helpers import build_model, rel_err
from oracle import ccvpe_oracle as orc

pytestmark = pytest.mark.gpu
TOL = 2e-4


def _gen(seed):
    return torch.Generator().manual_seed(seed)


def _cl(t, dtype, dev):
    return t.permute(0, 2, 3, 1).contiguous().to(dev, dtype)


def _wgrad_call(a0, a1, geom, g2d, N, row_scale, backend=cabi.BACKEND_SIMT):
    B, Hin, Win, Hout, Wout, stride, k, pad = geom
    d = cabi.WgradDesc()
    c0 = a0.shape[-1]
    c1 = a1.shape[-1] if a1 is not None else 0
    d.a0, d.a1 = a0.data_ptr(), (a1.data_ptr() if a1 is not None else None)
    d.c0, d.c1, d.ld0, d.ld1 = c0, c1, a0.stride(-2), (a1.stride(-2) if a1 is not None else 0)
    d.B, d.Hin, d.Win, d.Hout, d.Wout = B, Hin, Win, Hout, Wout
    d.stride, d.kh, d.kw, d.pad = stride, k, k, pad
    d.g, d.N, d.ldg = g2d.data_ptr(), N, g2d.stride(-2)
    d.g_row_scale = row_scale.data_ptr() if row_scale is not None else None
    d.dtype = cabi.dtype_code(a0.dtype)
    out = torch.empty((k * k, c0 + c1, N), dtype=torch.float32, device=a0.device)
    d.out = out.data_ptr()
    d.backend = backend
    n_ws = cabi.wgrad_workspace_elems(d)
    ws = torch.empty(n_ws, dtype=torch.float32, device=a0.device)
    d.workspace, d.workspace_elems = ws.data_ptr(), n_ws
    cabi.wgrad(d)
    torch.cuda.synchronize()
    return out


# ---------------------------------------------------------------------------------------------------------------
# weight gradients
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,c0,c1,cout,H,dtype,tol", [
    (2, 40, 24, 32, 16, torch.float32, TOL), (1, 16, 0, 16, 40, torch.float32, TOL), (3, 80, 112, 72, 8, torch.float32, TOL),
    (2, 16, 0, 8, 33, torch.float32, TOL), (2, 40, 24, 32, 16, torch.bfloat16, TOL),
])
def test_wgrad_conv3x3(cuda_device, B, c0, c1, cout, H, dtype, tol):
    """dW of a 3x3 pad-1 conv over two K-concatenated sources vs autograd (inputs rounded to `dtype`, fp32 math)."""
    g = _gen(31)
    dev = cuda_device
    x0 = torch.randn(B, c0, H, H, generator=g).to(dtype).float()
    x1 = torch.randn(B, c1, H, H, generator=g).to(dtype).float() if c1 else None
    dy = torch.randn(B, cout, H, H, generator=g).to(dtype).float()
    w = torch.zeros(cout, c0 + c1, 3, 3, requires_grad=True)
    xin = torch.cat([x0, x1], dim=1) if c1 else x0
    (F.conv2d(xin, w, padding=1) * dy).sum().backward()
    out = _wgrad_call(_cl(x0, dtype, dev), _cl(x1, dtype, dev) if c1 else None, (B, H, H, H, H, 1, 3, 1),
                      _cl(dy, dtype, dev).view(B * H * H, cout), cout, None)
    got = out.view(3, 3, c0 + c1, cout).permute(3, 2, 0, 1)
    assert rel_err(got, w.grad) < tol


@pytest.mark.parametrize("B,c0,c1,cout,H,W", [
    (2, 64, 0, 128, 16, 16),      # one c tile pair, one n tile, 8 pixel steps
    (3, 40, 24, 72, 8, 32),       # two sources with channel tails (40 = 32 + 8, 24), N tail (72 of 128), W = 32 boxes
    (1, 16, 0, 16, 64, 64),       # finest-level shape: 16 -> 16, split-K over pixels
    (2, 320, 112, 320, 32, 32),   # level-5 conv_a of the decoder: 14 c tiles x 3 n tiles
    (1, 16, 0, 8, 128, 128),      # padded 1-channel gradient (N = 8), W = 128: two boxes per row
])
def test_wgrad_conv3x3_tcgen05(cuda_device, B, c0, c1, cout, H, W):
    """Tensor-core weight gradient (MN-major UMMA operands, pixels on K) vs autograd on the same bf16-rounded inputs;
    fp32 accumulation on both sides: 1e-3 of max|ref|."""
    g = _gen(39)
    dev = cuda_device
    dt = torch.bfloat16
    x0 = torch.randn(B, c0, H, W, generator=g).to(dt).float()
    x1 = torch.randn(B, c1, H, W, generator=g).to(dt).float() if c1 else None
    dy = torch.randn(B, cout, H, W, generator=g).to(dt).float()
    w = torch.zeros(cout, c0 + c1, 3, 3, requires_grad=True)
    xin = torch.cat([x0, x1], dim=1) if c1 else x0
    (F.conv2d(xin, w, padding=1) * dy).sum().backward()
    out = _wgrad_call(_cl(x0, dt, dev), _cl(x1, dt, dev) if c1 else None, (B, H, W, H, W, 1, 3, 1),
                      _cl(dy, dt, dev).view(B * H * W, cout), cout, None, backend=cabi.BACKEND_TCGEN05)
    got = out.view(3, 3, c0 + c1, cout).permute(3, 2, 0, 1)
    assert rel_err(got, w.grad) < 1e-3, rel_err(got, w.grad)


@pytest.mark.parametrize("B,cin,cout,H,with_scale", [(2, 40, 16, 12, True), (1, 160, 80, 8, False), (2, 32, 24, 5, True),
                                                       (1, 64, 1024, 4, True)])      # deconv6: 1024 output channels
def test_wgrad_deconv_and_colsum(cuda_device, B, cin, cout, H, with_scale):
    """ConvTranspose2d(k2, s2) over [max | x * inv]: weight rows 1.. from ccvpe_wgrad (A = dY as a k2 s2 image, G = x with
    the F.normalize row scale), row 0 (max channel) and the bias from ccvpe_colsum -- vs autograd."""
    g = _gen(32)
    dev = cuda_device
    x = torch.randn(B, cin, H, H, generator=g)
    mx = torch.randn(B, 1, H, H, generator=g)
    inv = torch.rand(B, 1, H, H, generator=g) + 0.5 if with_scale else torch.ones(B, 1, H, H)
    dy = torch.randn(B, cout, 2 * H, 2 * H, generator=g)
    w = torch.zeros(cin + 1, cout, 2, 2, requires_grad=True)
    bias = torch.zeros(cout, requires_grad=True)
    (F.conv_transpose2d(torch.cat([mx, x * inv], dim=1), w, bias, stride=2) * dy).sum().backward()
    # dY embedded as the first `cout` channels of a wider tensor (as the conv_a data gradient hands it over)
    wide = torch.randn(B, 2 * H, 2 * H, cout + 24, generator=g).to(dev)
    wide[..., :cout] = dy.permute(0, 2, 3, 1).to(dev)
    d_up = wide[..., :cout]
    out = _wgrad_call(d_up, None, (B, 2 * H, 2 * H, H, H, 2, 2, 0), _cl(x, torch.float32, dev).view(B * H * H, cin), cin,
                      inv.reshape(-1).to(dev) if with_scale else None)
    got_rows = out.view(2, 2, cout, cin).permute(3, 2, 0, 1)
    assert rel_err(got_rows, w.grad[1:]) < TOL
    r1 = torch.empty(4, cout, device=dev)
    cabi.colsum(wide, cout, r1, w=mx.reshape(B, H, H).contiguous().to(dev), s=2)
    assert rel_err(r1.view(2, 2, cout).permute(2, 0, 1), w.grad[0]) < TOL
    bsum = torch.empty(1, cout, device=dev)
    cabi.colsum(wide, cout, bsum)
    assert rel_err(bsum[0], bias.grad) < TOL


@pytest.mark.parametrize("B,cin,cout,H,W", [(2, 64, 32, 8, 8), (3, 1280, 1024, 8, 8), (1, 40, 16, 64, 64),
                                             (2, 80, 40, 32, 128), (2, 1280, 72, 8, 8)])
def test_wgrad_k2s2_tcgen05(cuda_device, B, cin, cout, H, W):
    """Weight gradient of a k2 s2 transposed conv on tensor cores: A = dY gathered with stride 2 through a 5-D tensor map
    (four taps), G = the layer input; vs autograd on the same bf16-rounded inputs (1e-3 of max|ref|)."""
    g = _gen(40)
    dev = cuda_device
    dt = torch.bfloat16
    x = torch.randn(B, cin, H, W, generator=g).to(dt).float()
    dy = torch.randn(B, cout, 2 * H, 2 * W, generator=g).to(dt).float()
    w = torch.zeros(cin, cout, 2, 2, requires_grad=True)
    (F.conv_transpose2d(x, w, stride=2) * dy).sum().backward()
    out = _wgrad_call(_cl(dy, dt, dev), None, (B, 2 * H, 2 * W, H, W, 2, 2, 0), _cl(x, dt, dev).view(B * H * W, cin), cin, None,
                      backend=cabi.BACKEND_TCGEN05)
    got = out.view(2, 2, cout, cin).permute(3, 2, 0, 1)
    assert rel_err(got, w.grad) < 1e-3, rel_err(got, w.grad)


def test_scale_rows(cuda_device):
    g = _gen(44)
    x = torch.randn(3, 5, 7, 40, generator=g).to(cuda_device, torch.bfloat16)
    sc = torch.rand(3, 5, 7, generator=g).to(cuda_device)
    y = torch.empty_like(x)
    cabi.scale_rows(x, sc, y)
    assert torch.equal(y, (x.float() * sc[..., None]).to(torch.bfloat16))


def test_wgrad_cell(cuda_device):
    """Linear(5120 -> D) over 2x2 cells == conv k2 s2: dW vs autograd of the oracle's cell loop."""
    g = _gen(33)
    dev = cuda_device
    B, Cc, D = 2, 64, 48
    fs = torch.randn(B, Cc, 16, 16, generator=g)
    dy = torch.randn(B, D, 8, 8, generator=g)
    w = torch.zeros(D, Cc * 4, requires_grad=True)
    b = torch.zeros(D, requires_grad=True)
    (orc.sat_cell_descriptors(fs, w, b) * dy).sum().backward()
    out = _wgrad_call(_cl(fs, torch.float32, dev), None, (B, 16, 16, 8, 8, 2, 2, 0),
                      _cl(dy, torch.float32, dev).view(B * 64, D), D, None)
    got = out.view(2, 2, Cc, D).permute(3, 2, 0, 1).reshape(D, -1)
    assert rel_err(got, w.grad) < TOL


# ---------------------------------------------------------------------------------------------------------------
# pointwise helpers
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_relu_bwd_and_layout_helpers(cuda_device, dtype):
    g = _gen(34)
    dev = cuda_device
    h = torch.randn(2, 9, 7, 16, generator=g).to(dev, dtype)
    dh = torch.randn(2, 9, 7, 16, generator=g).to(dev, dtype)
    ref = torch.where(h > 0, dh, torch.zeros_like(dh))
    cabi.relu_bwd(dh, h)
    assert torch.equal(dh, ref)
    src = torch.randn(3, 2, 10, 11, generator=g).to(dev)
    dst = torch.full((3, 10, 11, 8), 5.0, device=dev, dtype=dtype)
    cabi.planar_to_cl(src, dst)
    assert torch.equal(dst[..., :2].float(), src.permute(0, 2, 3, 1).to(dtype).float()) and bool((dst[..., 2:] == 0).all())
    back = torch.empty(3, 2, 10, 11, device=dev)
    cabi.cl_to_planar(dst, 2, back)
    assert torch.equal(back, src.to(dtype).float())


def test_ori_normalize_bwd(cuda_device):
    g = _gen(35)
    dev = cuda_device
    v = torch.randn(2, 2, 16, 16, generator=g)
    v[0, :, 3, 4] = 0                                  # degenerate pixel: F.normalize clamps at eps
    d_ori = torch.randn(2, 2, 16, 16, generator=g)
    vv = v.clone().requires_grad_(True)
    (F.normalize(vv, p=2, dim=1) * d_ori).sum().backward()
    v_cl = torch.zeros(2, 16, 16, 2, device=dev)
    v_cl.copy_(v.permute(0, 2, 3, 1))
    dv = torch.full((2, 16, 16, 8), 3.0, device=dev)
    cabi.ori_normalize_bwd(v_cl, d_ori.to(dev), dv)
    mask = torch.ones(2, 16, 16, dtype=torch.bool)
    mask[0, 3, 4] = False                               # (the degenerate pixel's gradient is du / eps on both sides: huge)
    got = dv[..., :2].permute(0, 3, 1, 2).cpu()
    m4 = mask[:, None].expand(2, 2, 16, 16)
    assert rel_err(got[m4], vv.grad[m4]) < TOL
    assert bool((dv[..., 2:] == 0).all())


# ---------------------------------------------------------------------------------------------------------------
# matching + F.normalize backward
# ---------------------------------------------------------------------------------------------------------------
MATCH_BWD_CASES = [
    # name,        B, C,   L,   H,  rolls,                stride, centred
    ("vigor_l3",   2, 320, 320, 8,  list(range(20)),      16, False),
    ("vigor_l6",   1, 40,  40,  20, list(range(20)),      2,  False),     # HW = 400: ragged last tile
    ("fov180_l2",  2, 640, 320, 6,  list(range(-4, 5)),   32, False),
    ("kitti_l2",   1, 512, 256, 8,  list(range(16)),      64, False),     # shifts wrap past C
    ("oxford_l4",  2, 160, 28,  12, list(range(20)),      8,  True),
]


@pytest.mark.parametrize("case", MATCH_BWD_CASES, ids=[c[0] for c in MATCH_BWD_CASES])
@pytest.mark.parametrize("dtype,tol", [(torch.float32, TOL), (torch.bfloat16, 2e-2)])
def test_match_level_bwd(cuda_device, case, dtype, tol):
    """dx and dg of  scores = match(x, g);  mx = max_i scores;  xhat = F.normalize(x)  for incoming d_scores, d_max, d_xhat
    (two d_xhat contributions and a channels-last d_scores part, as at the bottleneck level) vs autograd of the oracle."""
    name, B, C, L, H, rolls, stride, centred = case
    g = _gen(36)
    dev = cuda_device
    R = len(rolls)
    x = torch.randn(B, C, H, H, generator=g).to(dtype).float()
    gd = torch.randn(B, L, generator=g)
    sel = [i for i in range(R) if i % 3 != 1]
    d_scores = torch.randn(B, R, H, H, generator=g)
    d_scores_cl = torch.randn(B, R, H, H, generator=g).to(dtype).float()
    d_max = torch.randn(B, 1, H, H, generator=g).to(dtype).float()
    d_xhat = torch.randn(B, C, H, H, generator=g).to(dtype).float()
    d_xhat2 = torch.randn(B, C, H, H, generator=g).to(dtype).float()
    xr = x.clone().requires_grad_(True)
    gr = gd.clone().requires_grad_(True)
    s = orc.match_level(xr, gr, rolls, stride, centred)
    loss = (s * (d_scores + d_scores_cl)).sum() + (s[:, sel].max(dim=1, keepdim=True)[0] * d_max).sum() \
        + (orc.l2_normalize(xr) * (d_xhat + d_xhat2)).sum()
    loss.backward()
    offset = int(C / 2 - L / 2) if centred else 0
    shifts = [i * stride for i in rolls]
    mask = sum(1 << i for i in sel)
    x_cl = _cl(x, dtype, dev)
    # the forward scores the backward consumes come from the library's own forward
    scores = torch.empty(B, R, H, H, device=dev)
    scratch = torch.empty(cabi.match_scratch_elems(B, C, R), device=dev)
    cabi.match_level(x_cl, gd.to(dev), offset, shifts, mask, scores=scores, scratch=scratch, backend=cabi.BACKEND_SIMT)
    # gradient operands laid out as the decoder hands them over: [d_xhat | d_max | pad] and [d_scores_cl pad 32 | d_xhat2]
    M = B * H * H
    a = torch.zeros(M, C + 8, device=dev, dtype=dtype)
    a[:, :C] = _cl(d_xhat, dtype, dev).view(M, C)
    a[:, C] = _cl(d_max, dtype, dev).view(M)
    b2 = torch.zeros(M, 32 + C, device=dev, dtype=dtype)
    b2[:, :R] = _cl(d_scores_cl, dtype, dev).view(M, R)
    b2[:, 32:] = _cl(d_xhat2, dtype, dev).view(M, C)
    dx = torch.empty_like(x_cl)
    dg = torch.empty(B, L, device=dev)
    cabi.match_level_bwd(x_cl, gd.to(dev), offset, shifts, mask, scores, d_scores.to(dev), b2, a[:, C:], a, b2[:, 32:], dx, dg)
    torch.cuda.synchronize()
    assert rel_err(dx.permute(0, 3, 1, 2).float(), xr.grad) < tol, rel_err(dx.permute(0, 3, 1, 2).float(), xr.grad)
    assert rel_err(dg, gr.grad) < tol, rel_err(dg, gr.grad)


# ---------------------------------------------------------------------------------------------------------------
# losses
# ---------------------------------------------------------------------------------------------------------------
def test_losses_match_oracle(cuda_device):
    g = _gen(37)
    dev = cuda_device
    B, n = 3, 20 * 16 * 16
    s = (torch.rand(B, n, generator=g) * 2 - 1)
    lab = torch.rand(B, n, generator=g) * (torch.rand(B, n, generator=g) > 0.97)
    sr = s.clone().requires_grad_(True)
    ref = orc.infonce_loss(sr, lab)
    ref.backward()
    sd = s.to(dev).requires_grad_(True)
    got = losses.infoNCELoss(sd, lab.to(dev))
    (3.0 * got).backward()
    assert abs(got.item() - ref.item()) < 1e-5 * abs(ref.item())
    assert rel_err(sd.grad / 3.0, sr.grad) < TOL
    # cross entropy on 512x512 logits with a normalised Gaussian label map
    gt, gwo, gor = synthetic_ground_truth(2, seed=3)
    lg = torch.randn(2, 512 * 512, generator=g) * 3
    lab2 = gt.flatten(1) / gt.flatten(1).sum(dim=1, keepdim=True)
    lr = lg.clone().requires_grad_(True)
    ref = orc.cross_entropy_loss(lr, lab2)
    ref.backward()
    ld = lg.to(dev).requires_grad_(True)
    got = losses.cross_entropy_loss(ld, lab2.to(dev))
    got.backward()
    assert abs(got.item() - ref.item()) < 1e-5 * abs(ref.item())
    assert rel_err(ld.grad, lr.grad) < TOL
    # orientation loss
    o = F.normalize(torch.randn(2, 2, 512, 512, generator=g), dim=1)
    orr = o.clone().requires_grad_(True)
    ref = orc.orientation_loss(orr, gor, gt)
    ref.backward()
    od = o.to(dev).requires_grad_(True)
    got = losses.orientation_loss(od, gor.to(dev), gt.to(dev))
    got.backward()
    assert abs(got.item() - ref.item()) < 1e-5 * abs(ref.item())
    assert rel_err(od.grad, orr.grad) < TOL


# ---------------------------------------------------------------------------------------------------------------
# ground descriptor heads
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,H,W,cs", [(2, 10, 20, (64, 32, 16, 8, 4, 2)), (1, 8, 32, (16, 8, 4, 2, 1, 1)), (2, 4, 7, (32, 1))])
def test_grd_descriptors_bwd(cuda_device, B, H, W, cs):
    g = _gen(38)
    dev = cuda_device
    K = 1280
    feat = torch.randn(B, K, H, W, generator=g)
    fr = feat.clone().requires_grad_(True)
    heads, refs, dgs = [], [], []
    total = 0
    for c in cs:
        w1 = (torch.randn(c, K, 1, 1, generator=g) * 0.05).requires_grad_(True)
        b1 = torch.randn(c, generator=g).requires_grad_(True)
        w2 = torch.randn(1, H, 1, 1, generator=g).requires_grad_(True)
        b2 = torch.randn(1, generator=g).requires_grad_(True)
        dg = torch.randn(B, W * c, generator=g)
        total = total + (orc.grd_descriptor(fr, w1, b1, w2, b2) * dg).sum()
        refs.append((w1, b1, w2, b2))
        heads.append((w1.detach().reshape(c, K).to(dev), b1.detach().to(dev), w2.detach().reshape(H).to(dev),
                      b2.detach().to(dev)))
        dgs.append(dg.to(dev))
    total.backward()
    dfeat = torch.empty(B, K, H, W, device=dev)
    dw1 = [torch.empty_like(h[0]) for h in heads]
    db1 = [torch.empty_like(h[1]) for h in heads]
    dw2 = [torch.empty_like(h[2]) for h in heads]
    db2 = [torch.empty_like(h[3]) for h in heads]
    cabi.grd_descriptors_bwd(feat.to(dev), heads, dgs, dfeat, dw1, db1, dw2, db2)
    torch.cuda.synchronize()
    assert rel_err(dfeat, fr.grad) < TOL
    for l, (w1, b1, w2, b2) in enumerate(refs):
        assert rel_err(dw1[l], w1.grad.reshape(dw1[l].shape)) < TOL, l
        assert rel_err(db1[l], b1.grad) < TOL, l
        assert rel_err(dw2[l], w2.grad.reshape(-1)) < TOL, l
        assert rel_err(db2[l], b2.grad) < TOL, l


# ---------------------------------------------------------------------------------------------------------------
# the whole training step
# ---------------------------------------------------------------------------------------------------------------
def _oracle_step(model, grd, sat, gts, train_mode, seed, variant="vigor"):
    """Loss and gradients of the reference training step (train_VIGOR.py:120-150) through autograd of the CPU oracle."""
    ref = copy.deepcopy(model).cpu()
    ref.train(train_mode)
    for p_ in ref.parameters():
        p_.requires_grad_(True)
    params = dict(ref.named_parameters())
    torch.manual_seed(seed)
    out = orc.forward_full(variant, params, ref.grd_efficientnet, ref.sat_efficientnet, grd, sat)
    loss = orc.training_loss(out, *gts)
    loss.backward()
    return loss.item(), {k: (v.grad.clone() if v.grad is not None else None) for k, v in params.items()}, [t.detach() for t in out]


@pytest.mark.parametrize("train_mode", [False, True])
def test_training_step_fp32_matches_oracle_autograd(cuda_device, train_mode, monkeypatch):
    """CVM_VIGOR, batch 2, fp32: forward outputs, the reference's combined loss and the gradient of EVERY parameter (98 head /
    decoder tensors through this library's backward kernels, and the encoder tensors through PyTorch autograd fed by the
    library's data gradients) against autograd through the CPU oracle.  train_mode=True exercises train-mode BatchNorm
    (batch statistics); stochastic depth is switched off for the comparison because the CPU and CUDA generators draw
    different masks from the same seed.  Loss within 1e-4 relative; gradients within 2e-3 of max|ref| per tensor."""
    from ccvpe_b200 import efficientnet
    from ccvpe_b200.synthetic import synthetic_pair
    monkeypatch.setattr(efficientnet, "_DROP_CONNECT", 0.0)
    model = build_model("vigor", None, True, 31)
    grd, sat = synthetic_pair(2, (320, 640), seed=61)
    gts = synthetic_ground_truth(2, seed=5)
    ref_loss, ref_grads, ref_out = _oracle_step(model, grd, sat, gts, train_mode, seed=123)
    dev = cuda_device
    gpu = copy.deepcopy(model).to(dev)
    gpu.train(train_mode)
    for p_ in gpu.parameters():
        p_.requires_grad_(True)
    cabi.reset_launch_count()
    torch.manual_seed(123)
    out = gpu(grd.to(dev), sat.to(dev))
    loss = losses.training_loss(out, *[t.to(dev) for t in gts])
    loss.backward()
    torch.cuda.synchronize()
    assert cabi.launch_count() > 150                                   # forward + backward kernels really ran
    for i, (a, b) in enumerate(zip(out, ref_out)):
        if i != 2:
            assert rel_err(a, b) < 1e-3, i
    assert abs(loss.item() - ref_loss) < 1e-4 * abs(ref_loss), (loss.item(), ref_loss)
    # Tolerance per tensor: 2e-3 of max|ref| of that tensor, plus an absolute floor of 2e-6 of the largest gradient entry of
    # the model for the tensors whose gradient is ZERO in exact arithmetic and pure summation noise on both sides: the bias of
    # the logits conv (sum over pixels of softmax - labels = 1 - 1) and, in train mode, every bias that is followed by a
    # batch-statistics BatchNorm (which removes any per-channel constant).
    gmax = max(float(v.abs().max()) for v in ref_grads.values() if v is not None)
    bad = {}
    for k, p_ in gpu.named_parameters():
        rg = ref_grads[k]
        if rg is None:
            assert p_.grad is None or float(p_.grad.abs().max()) == 0.0, k       # _fc.* are never used (models.py:151,166)
            continue
        assert p_.grad is not None, k
        err = float((p_.grad.detach().cpu().double() - rg.double()).abs().max())
        scale = float(rg.abs().max())
        if not err <= 2e-3 * scale + 2e-6 * gmax:
            bad[k] = (err, scale, gmax)
    assert not bad, sorted(bad.items(), key=lambda kv: -kv[1][0] / (kv[1][1] + 1e-30))[:10]


@pytest.mark.parametrize("variant,shape_key,n_bins", [("kitti", "kitti", 16), ("oxford", "oxford", 20)])
def test_training_step_fp32_other_classes(cuda_device, variant, shape_key, n_bins):
    """The same step for CVM_KITTI (16 orientations, windowed matching with wrap-around, reference train_KITTI.py:121-155) and
    CVM_OxfordRobotCar (centred windows, train_OxfordRobotCar.py:100-136), batch 1, eval-mode BatchNorm: loss and every
    gradient vs autograd through the oracle (same tolerances as the VIGOR test)."""
    from ccvpe_b200.synthetic import GROUND_SHAPES, synthetic_pair
    model = build_model(variant, None, None, 35)
    grd, sat = synthetic_pair(1, GROUND_SHAPES[shape_key], seed=65)
    gts = synthetic_ground_truth(1, seed=9, n_bins=n_bins)
    ref_loss, ref_grads, ref_out = _oracle_step(model, grd, sat, gts, False, seed=3, variant=variant)
    dev = cuda_device
    gpu = copy.deepcopy(model).to(dev).eval()
    for p_ in gpu.parameters():
        p_.requires_grad_(True)
    out = gpu(grd.to(dev), sat.to(dev))
    loss = losses.training_loss(out, *[t.to(dev) for t in gts])
    loss.backward()
    torch.cuda.synchronize()
    assert abs(loss.item() - ref_loss) < 1e-4 * abs(ref_loss), (loss.item(), ref_loss)
    gmax = max(float(v.abs().max()) for v in ref_grads.values() if v is not None)
    bad = {}
    for k, p_ in gpu.named_parameters():
        rg = ref_grads[k]
        if rg is None:
            continue
        err = float((p_.grad.detach().cpu().double() - rg.double()).abs().max())
        if not err <= 2e-3 * float(rg.abs().max()) + 2e-6 * gmax:
            bad[k] = (err, float(rg.abs().max()))
    assert not bad, sorted(bad.items(), key=lambda kv: -kv[1][0])[:10]


def test_losses_and_model_refuse_cpu_tensors():
    """No CPU fallback on the training path either."""
    with pytest.raises(cabi.CcvpeError):
        losses.infoNCELoss(torch.zeros(1, 8), torch.zeros(1, 8))
    with pytest.raises(cabi.CcvpeError):
        losses.cross_entropy_loss(torch.zeros(1, 8), torch.zeros(1, 8))
    with pytest.raises(cabi.CcvpeError):
        losses.orientation_loss(torch.zeros(1, 2, 4, 4), torch.zeros(1, 2, 4, 4), torch.zeros(1, 1, 4, 4))


def test_training_step_bf16_close_to_fp32_oracle(cuda_device):
    """bf16 training path (bf16-autocast encoders, bf16 activations, tcgen05 forward / data-gradient / weight-gradient GEMMs
    with fp32 accumulation, fp32 master weights): the loss within 2 % of the fp32 oracle's and every head / decoder gradient
    tensor with more than 1000 elements has cosine similarity >= 0.95 with it (measured: >= 0.96, all but one >= 0.98)."""
    from ccvpe_b200.synthetic import synthetic_pair
    model = build_model("vigor", None, True, 32)
    grd, sat = synthetic_pair(2, (320, 640), seed=62)
    gts = synthetic_ground_truth(2, seed=6)
    ref_loss, ref_grads, _ = _oracle_step(model, grd, sat, gts, False, seed=7)
    dev = cuda_device
    gpu = copy.deepcopy(model).to(dev).set_precision("bf16")
    gpu.eval()
    for p_ in gpu.parameters():
        p_.requires_grad_(True)
    out = gpu(grd.to(dev), sat.to(dev))
    loss = losses.training_loss(out, *[t.to(dev) for t in gts])
    loss.backward()
    torch.cuda.synchronize()
    assert abs(loss.item() - ref_loss) < 2e-2 * abs(ref_loss), (loss.item(), ref_loss)
    low = {}
    for k, p_ in gpu.named_parameters():
        rg = ref_grads[k]
        if rg is None or rg.numel() < 1000 or k.startswith(("grd_efficientnet", "sat_efficientnet")):
            continue
        cs = F.cosine_similarity(p_.grad.flatten().cpu().double(), rg.flatten().double(), dim=0).item()
        if cs < 0.95:
            low[k] = cs
    assert not low, low


def test_graphed_training_step_matches_eager(cuda_device):
    """training.GraphedTrainStep: the whole step (zero_grad, forward, losses, backward, Adam) captured in one CUDA graph gives
    the same losses as the same steps run eagerly from the same start (bf16 path, eval-mode BatchNorm so that both runs are
    deterministic functions of the weights)."""
    from ccvpe_b200.synthetic import synthetic_pair
    from ccvpe_b200.training import GraphedTrainStep
    dev = cuda_device
    batch = [t.to(dev) for t in synthetic_pair(2, (320, 640), seed=64)] + [t.to(dev) for t in synthetic_ground_truth(2, seed=8)]
    hist = {}
    for mode in ("eager", "graph"):
        model = build_model("vigor", None, True, 34).to(dev).set_precision("bf16").eval()
        for n_, p_ in model.named_parameters():
            p_.requires_grad_("._fc." not in n_)
        opt = torch.optim.Adam([p_ for p_ in model.parameters() if p_.requires_grad], lr=1e-4, fused=True, capturable=True)

        def fn(grd, sat, gt, gwo, gor):
            opt.zero_grad(set_to_none=False)
            loss = losses.training_loss(model(grd, sat), gt, gwo, gor)
            loss.backward()
            opt.step()
            return loss.detach()

        if mode == "eager":
            hist[mode] = [float(fn(*batch)) for _ in range(6)]
        else:
            # give .grad static storage before capture (zero_grad(set_to_none=False) keeps it)
            fn(*batch)                                               # step 1
            step = GraphedTrainStep(fn, batch, warmup=1)            # step 2 (warm-up); the capture itself executes nothing
            hist[mode] = [float("nan")] * 2 + [float(step(*batch)) for _ in range(4)]      # steps 3..6 are replays
    for a, b in zip(hist["eager"][2:], hist["graph"][2:]):
        assert abs(a - b) <= 1e-3 * abs(a), hist
    assert hist["eager"][-1] < hist["eager"][0]


def test_adam_steps_reduce_the_loss(cuda_device):
    """Five Adam(1e-4) steps on one fixed synthetic batch (train_VIGOR.py:101-150) through the CUDA path: the loss goes down
    and the weight cache follows the updated parameters."""
    from ccvpe_b200.synthetic import synthetic_pair
    dev = cuda_device
    model = build_model("vigor", None, True, 33).to(dev).train()
    for p_ in model.parameters():
        p_.requires_grad_(True)
    opt = torch.optim.Adam([p_ for p_ in model.parameters() if p_.requires_grad], lr=1e-4, betas=(0.9, 0.999))
    grd, sat = (t.to(dev) for t in synthetic_pair(2, (320, 640), seed=63))
    gts = [t.to(dev) for t in synthetic_ground_truth(2, seed=7)]
    hist = []
    for _ in range(5):
        opt.zero_grad()
        loss = losses.training_loss(model(grd, sat), *gts)
        loss.backward()
        opt.step()
        hist.append(loss.item())
    assert all(math.isfinite(v) for v in hist) and hist[-1] < hist[0], hist
