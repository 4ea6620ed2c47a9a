"""Summarises the parity numbers recorded by the GPU tests (gpurun_out/parity_*.json) into profiles/<tag>_parity.md.
usage: python scripts/summarize_parity.py [tag]   (default r03)"""
import glob
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG = sys.argv[1] if len(sys.argv) > 1 else "r03"
out = ["# Parity numbers measured on the B200 (round 2; tests/test_gpu_forward.py)", ""]
out += ["## fp32 path: orientation field v/|v| against the CPU oracle, and the noise floor", "",
        "`ours` = libccvpe_b200 fp32 path; `floor` = the ORACLE ITSELF executed through cuDNN/cuBLAS fp32 (TF32 off) on the same GPU.",
        "Columns: plain max |ori - ref| over all pixels; fraction of pixels above 1e-3; max error over pixels with |v| >= 2 % of max|v|;",
        "logits max rel err.", "",
        "| config | well-conditioned px | ours max | ours frac>1e-3 | ours max (well-cond.) | floor max | floor frac>1e-3 | floor max (well-cond.) | logits ours / floor |",
        "|---|---|---|---|---|---|---|---|---|"]
for f in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "parity_ori_fp32_*.json"))):
    d = json.load(open(f))
    o, fl = d["ours"], d["oracle_cudnn_fp32"]
    out.append("| %s | %.4f | %.2e | %.1e | %.2e | %.2e | %.1e | %.2e | %.1e / %.1e |" % (
        d["config"], d["well_conditioned_fraction"], o["plain_max_err"], o["frac_pixels_above_1e-3"], o["max_err_well_conditioned"],
        fl["plain_max_err"], fl["frac_pixels_above_1e-3"], fl["max_err_well_conditioned"], d["logits_rel_err"]["ours"],
        d["logits_rel_err"]["oracle_cudnn_fp32"]))
out += ["", "## bf16 path at the benchmarked sizes against the fp32 CPU oracle", "",
        "max = max|err| / max|ref|, rms = rms(err) / rms(ref), per tensor.", ""]
for f in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "parity_bf16_*.json"))):
    d = json.load(open(f))
    out.append("### %s, batch %d" % (d["variant"], d["batch"]))
    out.append("")
    out.append("| tensor | max | rms |")
    out.append("|---|---|---|")
    for k in ["logits", "heatmap"] + ["scores%d" % i for i in range(1, 7)]:
        out.append("| %s | %.4f | %.4f |" % (k, d[k]["max_rel"], d[k]["rms_rel"]))
    o = d["ori"]
    out.append("")
    out.append("orientation field (angle error where |v| > 1 %% of max|v|, %.2f %% of pixels): median %.3f deg, p95 %.3f deg, max %.1f deg"
               % (100 * o["pixels_compared_fraction"], o["angle_deg_median"], o["angle_deg_p95"], o["angle_deg_max"]))
    a = d["argmax"]
    out.append("argmax: %d / %d pairs equal to the oracle's; %d / %d of the pairs whose top-2 logit gap exceeds 10x the rms logit error (%.4f)"
               % (a["agree"], a["pairs"], a["agree_decided"], a["decided_pairs"], a["logit_rms_err"]))
    out.append("")
open(os.path.join(ROOT, "profiles", TAG + "_parity.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
