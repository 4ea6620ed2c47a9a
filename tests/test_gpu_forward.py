"""GPU end-to-end parity: the drop-in models (PyTorch encoders + CUDA decoder path through the C ABI) against
(1) the CPU oracle run live on the same seeded weights/inputs and (2) the committed fixtures generated from the
unmodified reference.  fp32 tolerance: 1e-3 of max|ref| per tensor (BASELINE.json); argmax / pixel locations bit-exact.
"""
import os

import numpy as np
import pytest
import torch

from ccvpe_b200 import cabi
from helpers import (GOLDEN_CONFIGS, OUT_NAMES, build_model, check_against_golden, config_inputs, oracle_forward,
                     ori_field_err, rel_err)
from oracle import ccvpe_oracle as orc

pytestmark = pytest.mark.gpu
GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FP32_TOL = 1e-3


@pytest.mark.parametrize("name", sorted(GOLDEN_CONFIGS))
def test_forward_fp32_matches_oracle_and_reference_fixture(cuda_device, name):
    variant, shape_key, noise, circular, batch, wseed, iseed = GOLDEN_CONFIGS[name]
    model = build_model(variant, noise, circular, wseed)
    grd, sat = config_inputs(name)
    inter = {}
    ref = oracle_forward(model, variant, noise, grd, sat, inter)
    gpu_model = model.to(cuda_device)
    cabi.reset_launch_count()
    with torch.no_grad():
        out = gpu_model(grd.to(cuda_device), sat.to(cuda_device))
    torch.cuda.synchronize()
    assert cabi.launch_count() > 40                      # the CUDA path really ran (no silent fallback)
    assert len(out) == 9
    for n, a, b in zip(OUT_NAMES, out, ref):
        assert a.shape == b.shape, n
        assert a.dtype == torch.float32 and a.is_contiguous(), n
        # the unit orientation field is judged on the pre-normalisation field (see helpers.ori_field_err); its plain
        # max error is additionally bounded by 1e-2 (it exceeds 1e-3 only at the few pixels where |v| is ~0)
        err = ori_field_err(a, b, inter["ori_raw"]) if n == "ori" else rel_err(a, b)
        assert err < FP32_TOL, "%s rel err %.3e" % (n, err)
    assert rel_err(out[2], ref[2]) < 1e-2
    # fixtures from the unmodified reference
    golden = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    check_against_golden(out, golden, tol=FP32_TOL, loose={"ori": 1e-2})
    # pose decode on the device == numpy decode of the reference outputs: indices bit-exact
    pose = {k: v.cpu().numpy() for k, v in gpu_model.decode_pose(out[1], out[2]).items()}
    assert pose["idx"].tolist() == golden["pose.idx"].tolist()
    assert pose["rc"].tolist() == golden["pose.rc"].tolist()
    assert pose["valid"].tolist() == golden["pose.valid"].tolist()
    np.testing.assert_allclose(pose["cs"], golden["pose.cs"], atol=2e-3)
    dang = np.abs(pose["angle"] - golden["pose.angle"])
    assert np.all(np.minimum(dang, 360 - dang) < 0.5)    # degrees; (cos, sin) within 1e-3 -> angle well within 0.5


def test_batch_independence_and_localize(cuda_device):
    """Pairs are independent: a batch of 3 equals three batches of 1, bit for bit."""
    model = build_model("vigor", None, True, 11).to(cuda_device)
    g = torch.Generator().manual_seed(5)
    grd = torch.randn(3, 3, 320, 640, generator=g).to(cuda_device)
    sat = torch.randn(3, 3, 512, 512, generator=g).to(cuda_device)
    with torch.no_grad():
        full = model(grd, sat)
        singles = [model(grd[i:i + 1], sat[i:i + 1]) for i in range(3)]
    for k in range(9):
        cat = torch.cat([s[k] for s in singles], dim=0)
        # cuDNN may pick batch-dependent encoder algorithms -> last-bit differences; the unit orientation field
        # amplifies them where the raw field is ~0 (see helpers.ori_field_err)
        assert rel_err(full[k], cat) < (1e-2 if OUT_NAMES[k] == "ori" else 1e-5), OUT_NAMES[k]
    pose = model.localize(grd, sat)
    assert pose["idx"].tolist() == [int(full[1][i].flatten().argmax()) for i in range(3)]


def test_no_cpu_fallback(cuda_device):
    model = build_model("vigor", None, True, 1)
    with pytest.raises(cabi.CcvpeError):
        model(torch.zeros(1, 3, 320, 640), torch.zeros(1, 3, 512, 512))


def test_weight_cache_tracks_parameter_updates(cuda_device):
    model = build_model("oxford", None, None, 2).to(cuda_device)
    grd = torch.randn(1, 3, 154, 231, device=cuda_device)
    sat = torch.randn(1, 3, 512, 512, device=cuda_device)
    with torch.no_grad():
        a = model(grd, sat)[0].clone()
        model.conv1[2].bias.add_(1.0)
        b = model(grd, sat)[0]
    assert torch.allclose(b, a + 1.0, atol=1e-4)


@pytest.mark.parametrize("name", ["kitti_b1", "oxford_b1", "vigor_prior72_fov180_b1", "vigor_prior180_fov360_b1"])
def test_bf16_forward_other_classes(cuda_device, name):
    """bf16 path on the classes whose matching is windowed (L < C: CUDA-core match kernel on bf16 maps) or prior-limited,
    and on the KITTI channel plan (2048-d cells, 88->128 widening conv).  Stated bf16 tolerance: max 1.5e-1 of max|ref|,
    rms 8e-2 of rms(ref) (windowed cosines over as few as 7-64 channels are noisier than the full-circle VIGOR ones)."""
    variant, shape_key, noise, circular, batch, wseed, iseed = GOLDEN_CONFIGS[name]
    model = build_model(variant, noise, circular, wseed)
    grd, sat = config_inputs(name)
    ref = oracle_forward(model, variant, noise, grd, sat)
    gpu_model = model.to(cuda_device).set_precision("bf16")
    with torch.no_grad():
        out = gpu_model(grd.to(cuda_device), sat.to(cuda_device))
    assert len(out) == 9
    for n, a, b in zip(OUT_NAMES, out, ref):
        assert a.shape == b.shape and a.dtype == torch.float32, n
        assert torch.isfinite(a).all(), n
        if n == "ori":
            continue
        assert rel_err(a, b) < 1.5e-1, "%s max rel err %.3e" % (n, rel_err(a, b))
        rms = ((a.cpu().double() - b.double()).pow(2).mean().sqrt() / b.double().pow(2).mean().sqrt()).item()
        assert rms < 8e-2, "%s rms rel err %.3e" % (n, rms)
    cosang = (out[2].cpu() * ref[2]).sum(dim=1)
    assert (cosang > 0.95).float().mean() > 0.95


def test_bf16_forward_within_stated_tolerance(cuda_device):
    """bf16 path (bf16 encoders + bf16 decoder kernels, fp32 accumulation).  Stated bf16 tolerance vs the fp32 oracle,
    per tensor: rms(err) <= 6e-2 * rms(ref) and max|err| <= 1.5e-1 * max|ref|; and the argmax must equal the fp32
    oracle's for every pair whose top-2 logit gap exceeds 10x the measured rms logit error.  (SURVEY section 8(c) measured
    0.4-1.5e-2 max-abs for the reference itself under bf16 autocast with its near-constant default-init outputs; the
    He-scaled test weights keep O(1) activations through 2 encoders + 12 convs, so single-pixel extremes are larger.)"""
    name = "vigor_fov360_b2"
    variant, shape_key, noise, circular, batch, wseed, iseed = GOLDEN_CONFIGS[name]
    model = build_model(variant, noise, circular, wseed)
    grd, sat = config_inputs(name)
    ref = oracle_forward(model, variant, noise, grd, sat)
    gpu_model = model.to(cuda_device).set_precision("bf16")
    with torch.no_grad():
        out = gpu_model(grd.to(cuda_device), sat.to(cuda_device))
    for n, a, b in zip(OUT_NAMES, out, ref):
        assert a.shape == b.shape and a.dtype == torch.float32
        if n == "ori":
            # unit vectors: compare where the un-normalised field is not tiny (direction is ill-conditioned there)
            continue
        err = rel_err(a, b)
        assert err < 1.5e-1, "%s max rel err %.3e" % (n, err)
        rms = ((a.cpu().double() - b.double()).pow(2).mean().sqrt() / b.double().pow(2).mean().sqrt()).item()
        assert rms < 6e-2, "%s rms rel err %.3e" % (n, rms)
    cosang = (out[2].cpu() * ref[2]).sum(dim=1)
    assert (cosang > 0.95).float().mean() > 0.97
    logit_rms = (out[0].cpu() - ref[0]).pow(2).mean().sqrt().item()
    top = torch.topk(ref[0], 2, dim=1).values
    checked = 0
    for b in range(ref[0].shape[0]):
        if (top[b, 0] - top[b, 1]).item() > 10 * logit_rms:
            assert int(out[0][b].argmax()) == int(ref[0][b].argmax())
            checked += 1
    assert checked >= 1


def test_bf16_path_is_bit_reproducible_and_graph_equals_eager(cuda_device):
    """The bf16 path has no order-dependent arithmetic (the squeeze-excite sums are accumulated in fixed point), so the
    same inputs give the same bits: eager twice, and a CUDA-graph replay (`set_cuda_graph(True)`) of the same kernels.
    Inputs are re-read on every replay, and replays are counted as launches."""
    from ccvpe_b200 import cabi, models
    from ccvpe_b200.synthetic import fill_deterministic, synthetic_pair
    model = models.CVM_VIGOR("cuda", True).eval()
    fill_deterministic(model.state_dict(), seed=5)
    model = model.to(cuda_device).set_precision("bf16")
    pairs = [tuple(t.to(cuda_device) for t in synthetic_pair(2, (320, 640), seed=s)) for s in (11, 12)]
    with torch.no_grad():
        eager = [[t.clone() for t in model(g, s)] for g, s in pairs]
        eager_again = [t.clone() for t in model(*pairs[0])]
        model.set_cuda_graph(True)
        first = [t.clone() for t in model(*pairs[0])]            # capture + replay
        cabi.reset_launch_count()
        second = [t.clone() for t in model(*pairs[1])]           # replay with new inputs
        launches = cabi.launch_count()
        again = [t.clone() for t in model(*pairs[0])]
    torch.cuda.synchronize()
    assert launches > 100, launches
    for i, name in enumerate(OUT_NAMES):
        assert torch.equal(eager_again[i], eager[0][i]), "eager run-to-run: " + name
        assert torch.equal(first[i], eager[0][i]), "graph vs eager: " + name
        assert torch.equal(second[i], eager[1][i]), "graph replay with new inputs: " + name
        assert torch.equal(again[i], eager[0][i]), "graph replay: " + name
    assert rel_err(second[0], first[0]) > 1e-1                   # different inputs -> different logits
    model.set_cuda_graph(False)


def _record(name, payload):
    """Measured parity numbers are also written to gpurun_out/parity_<name>.json (summarised under profiles/)."""
    import json
    out_dir = os.path.join(os.path.dirname(GOLDEN_DIR), "..", "gpurun_out")
    try:
        os.makedirs(out_dir, exist_ok=True)
        with open(os.path.join(out_dir, "parity_%s.json" % name), "w") as f:
            json.dump(payload, f, indent=1, sort_keys=True)
    except OSError:
        pass


def _oracle_on(device, model, variant, noise, grd, sat, inter=None):
    """The oracle's torch ops executed on `device` (cuDNN / cuBLAS fp32 with TF32 off on the GPU): the same reference
    arithmetic on a different backend -- its distance from the CPU oracle is the noise floor of any fp32 comparison."""
    import copy
    m = copy.deepcopy(model).to(device)
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    with torch.no_grad():
        return orc.forward_full(variant, sd, m.grd_efficientnet, m.sat_efficientnet, grd.to(device), sat.to(device), noise,
                                inter)


#: the unit orientation field v/|v| is compared with the plain 1e-3 bar wherever |v| >= ORI_COND * max|v|
ORI_COND = 2e-2


@pytest.mark.parametrize("name", sorted(GOLDEN_CONFIGS))
def test_orientation_field_parity_quantified(cuda_device, name):
    """north_star: "orientation fields within 1e-3 relative in fp32".  ori = v/|v| amplifies any absolute error e on the raw
    2-vector v to e/|v|, so two correct fp32 evaluations of the reference arithmetic differ by more than 1e-3 at pixels
    where |v| is tiny.  This test states the bar that IS met and quantifies the rest:
      (1) plain |ori - ori_ref| <= 1e-3 at every pixel with |v| >= ORI_COND * max|v| (the well-conditioned pixels);
      (2) the pixels above 1e-3 are a small fraction and no more numerous than for the ORACLE ITSELF run through cuDNN fp32
          on the same GPU (noise floor: same arithmetic, different summation order) -- within 3x + 1e-4;
      (3) the numbers (plain max error, fraction > 1e-3, floor) are recorded per config."""
    variant, shape_key, noise, circular, batch, wseed, iseed = GOLDEN_CONFIGS[name]
    model = build_model(variant, noise, circular, wseed)
    grd, sat = config_inputs(name)
    inter = {}
    ref = oracle_forward(model, variant, noise, grd, sat, inter)
    floor = _oracle_on(cuda_device, model, variant, noise, grd, sat)
    gpu_model = model.to(cuda_device)
    with torch.no_grad():
        out = gpu_model(grd.to(cuda_device), sat.to(cuda_device))
    raw = inter["ori_raw"].double()
    mag = raw.norm(dim=1, keepdim=True)
    well = (mag >= ORI_COND * mag.max()).expand_as(ref[2])
    rec = {"config": name, "ori_cond_threshold": ORI_COND, "well_conditioned_fraction": float(well.float().mean())}
    for tag, cand in (("ours", out[2]), ("oracle_cudnn_fp32", floor[2])):
        d = (cand.detach().cpu().double() - ref[2].double()).abs()
        rec[tag] = {"plain_max_err": float(d.max()), "frac_pixels_above_1e-3": float((d > 1e-3).float().mean()),
                    "max_err_well_conditioned": float(d[well].max()),
                    "raw_field_weighted_err": float((d * mag / raw.abs().max()).max())}
    rec["logits_rel_err"] = {"ours": rel_err(out[0], ref[0]), "oracle_cudnn_fp32": rel_err(floor[0], ref[0])}
    _record("ori_fp32_" + name, rec)
    assert rec["ours"]["max_err_well_conditioned"] <= 1e-3, rec
    assert rec["ours"]["frac_pixels_above_1e-3"] <= 3 * rec["oracle_cudnn_fp32"]["frac_pixels_above_1e-3"] + 1e-4, rec


def _chunked_oracle(model, variant, noise, grd, sat, chunk=4):
    outs, raws = [], []
    for i in range(0, grd.shape[0], chunk):
        inter = {}
        o = oracle_forward(model, variant, noise, grd[i:i + chunk], sat[i:i + chunk], inter)
        outs.append([t.clone() for t in o])
        raws.append(inter["ori_raw"].clone())
    return [torch.cat([o[k] for o in outs], dim=0) for k in range(9)], torch.cat(raws, dim=0)


#: stated bf16 tolerance of the full-size configurations (bf16 encoders + bf16 tcgen05 decoder vs the fp32 oracle), per tensor.
#: Measured on B200 (profiles/r02_parity.md): max <= 2.3e-2, rms <= 1.9e-2, orientation p95 <= 0.95 deg -- SURVEY section 8(c)'s
#: proposed 2e-2 * max|ref| with a 1.7x margin.
BF16_MAX_TOL = 4e-2        # max|err| / max|ref|
BF16_RMS_TOL = 3e-2        # rms(err) / rms(ref)
BF16_ORI_DEG = 3.0         # 95th percentile of the angular error of the orientation field where |v| > 1 % of max|v|


@pytest.mark.parametrize("variant,shape_key,batch,wseed", [("vigor", "vigor", 64, 21), ("kitti", "kitti", 32, 22)])
def test_bf16_full_size_batches_against_oracle(cuda_device, variant, shape_key, batch, wseed):
    """The BENCHMARKED configurations -- CVM_VIGOR batch 64 bf16 (BASELINE.json configs[1]) and CVM_KITTI batch 32
    (configs[2]) -- through the tcgen05 kernels, against the fp32 CPU oracle on the same weights and inputs (oracle
    evaluated in chunks of 4 pairs).  Per tensor: max and rms error; orientation field by ANGLE where the raw field is not
    tiny; argmax equality for every pair whose top-2 logit gap exceeds 10x the rms logit error.  Served from the model's
    CUDA-graph mode, exactly as bench.py times it."""
    from ccvpe_b200.synthetic import GROUND_SHAPES, synthetic_pair
    model = build_model(variant, None, True if variant == "vigor" else None, wseed)
    grd, sat = synthetic_pair(batch, GROUND_SHAPES[shape_key], seed=40 + wseed)
    ref, raw = _chunked_oracle(model, variant, None, grd, sat)
    gpu_model = model.to(cuda_device).set_precision("bf16").set_cuda_graph(True)
    cabi.reset_launch_count()
    with torch.no_grad():
        out = [t.clone() for t in gpu_model(grd.to(cuda_device), sat.to(cuda_device))]
    torch.cuda.synchronize()
    assert cabi.launch_count() > 100
    rec = {"variant": variant, "batch": batch}
    for n, a, b in zip(OUT_NAMES, out, ref):
        assert a.shape == b.shape and a.dtype == torch.float32 and torch.isfinite(a).all(), n
        if n == "ori":
            continue
        err = rel_err(a, b)
        rms = ((a.cpu().double() - b.double()).pow(2).mean().sqrt() / b.double().pow(2).mean().sqrt()).item()
        rec[n] = {"max_rel": err, "rms_rel": rms}
    mag = raw.double().norm(dim=1)
    keep = mag > 1e-2 * mag.max()
    cosang = (out[2].cpu().double() * ref[2].double()).sum(dim=1).clamp(-1, 1)
    ang = torch.rad2deg(torch.acos(cosang))[keep]
    rec["ori"] = {"angle_deg_median": float(ang.median()), "angle_deg_p95": float(ang.quantile(0.95)),
                  "angle_deg_max": float(ang.max()), "pixels_compared_fraction": float(keep.float().mean())}
    logit_rms = (out[0].cpu() - ref[0]).pow(2).mean().sqrt().item()
    top = torch.topk(ref[0], 2, dim=1).values
    gaps = (top[:, 0] - top[:, 1])
    decided = [b for b in range(batch) if gaps[b].item() > 10 * logit_rms]
    agree = sum(int(out[0][b].argmax()) == int(ref[0][b].argmax()) for b in range(batch))
    agree_decided = sum(int(out[0][b].argmax()) == int(ref[0][b].argmax()) for b in decided)
    rec["argmax"] = {"pairs": batch, "agree": agree, "decided_pairs": len(decided), "agree_decided": agree_decided,
                     "logit_rms_err": logit_rms}
    _record("bf16_%s_b%d" % (variant, batch), rec)
    for n in OUT_NAMES:
        if n != "ori":
            assert rec[n]["max_rel"] < BF16_MAX_TOL and rec[n]["rms_rel"] < BF16_RMS_TOL, (n, rec[n])
    assert rec["ori"]["angle_deg_p95"] < BF16_ORI_DEG, rec["ori"]
    assert agree_decided == len(decided) and len(decided) >= 1, rec["argmax"]


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_uint8_images_are_ingested_on_the_device(cuda_device, precision):
    """forward(uint8 images) == forward(ToTensor + ImageNet Normalize of the same pixels): bit-identical on the fp32 path
    (ccvpe_ingest_u8 reproduces torchvision's arithmetic), within bf16 noise on the bf16 plan (the stem kernel normalises
    with one FMA in its loads).  Odd batch size on purpose."""
    model = build_model("vigor", None, True, 12).to(cuda_device).set_precision(precision)
    g = torch.Generator().manual_seed(6)
    grd8 = torch.randint(0, 256, (3, 3, 320, 640), generator=g, dtype=torch.uint8)
    sat8 = torch.randint(0, 256, (3, 3, 512, 512), generator=g, dtype=torch.uint8)
    mean = torch.tensor(cabi.IMAGENET_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(cabi.IMAGENET_STD).view(1, 3, 1, 1)
    grd = ((grd8.float().div(255) - mean) / std).to(cuda_device)
    sat = ((sat8.float().div(255) - mean) / std).to(cuda_device)
    with torch.no_grad():
        a = model(grd8.to(cuda_device), sat8.to(cuda_device))
        b = model(grd, sat)
        pose = model.localize_u8(grd8.to(cuda_device), sat8.to(cuda_device))
    for i, n in enumerate(OUT_NAMES):
        if precision == "fp32":
            assert torch.equal(a[i], b[i]), n
        elif n != "ori":
            assert rel_err(a[i], b[i]) < 4e-2, (n, rel_err(a[i], b[i]))
    assert pose["idx"].tolist() == [int(a[1][i].flatten().argmax()) for i in range(3)]


def test_model_follows_its_tensors_device(cuda_device):
    """a model moved to cuda:1 while the current device is cuda:0 must launch on cuda:1 (kernel attributes,
    stream and tensor maps are per device)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    model = build_model("oxford", None, None, 2)
    grd, sat = config_inputs("oxford_b1")
    with torch.no_grad():
        a = model.to("cuda:0")(grd.to("cuda:0"), sat.to("cuda:0"))
        a = [t.cpu() for t in a]
        torch.cuda.set_device(0)
        m1 = model.to("cuda:1")
        b = m1(grd.to("cuda:1"), sat.to("cuda:1"))
        pose = m1.decode_pose(b[1], b[2])
        m1.set_precision("bf16")
        c = m1(grd.to("cuda:1"), sat.to("cuda:1"))
    assert all(t.device.index == 1 for t in b) and pose["idx"].device.index == 1 and c[0].device.index == 1
    for x, y in zip(a, b):
        assert rel_err(y, x) < 1e-5
    with pytest.raises(cabi.CcvpeError):                   # direct binding calls on a non-current device are refused loudly
        cabi.softmax_heatmap(b[0], torch.empty_like(b[0]), torch.empty(1024, device="cuda:1"))
