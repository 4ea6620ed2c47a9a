"""Collects the bench.py JSON lines recorded under gpurun_out/ into profiles/<tag>_bench_lines.md (profiles tool).
usage: python scripts/summarize_bench.py [tag] [single-GPU run prefix]   (defaults: r03 r03k)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG = sys.argv[1] if len(sys.argv) > 1 else "r03"
RUN = sys.argv[2] if len(sys.argv) > 2 else "r03k"
RUNS = [
    ("vigor_b64 (default workload), 1 GPU", RUN + "_bench.json"),
    ("kitti_b32, 1 GPU", RUN + "_bench_kitti_b32.json"),
    ("vigor_prior72_fov180, 1 GPU", RUN + "_bench_vigor_prior72_fov180.json"),
    ("vigor_prior72_fov108, 1 GPU", RUN + "_bench_vigor_prior72_fov108.json"),
    ("oxford_b1 (batch-1 sequential frames), 1 GPU", RUN + "_bench_oxford_b1.json"),
    ("train, B=8 bf16, one CUDA graph per step, 1 GPU", RUN + "_bench_train.json"),
    ("train, B=8 fp32 parity path (eager), 1 GPU", "r02f_bench_train_fp32.json"),
    ("2 GPUs, weak (64 pairs per GPU), final code", "r03q_n2_weak.json"),
    ("2 GPUs, strong (one batch of 64), final code", "r03q_n2_strong.json"),
    ("2 GPUs, weak (64 pairs per GPU), earlier in the round", "r02m_n2_weak.json"),
    ("2 GPUs, strong (one batch of 64), earlier in the round", "r02m_n2_strong.json"),
    ("2 GPUs, train (graph)", "r02m_n2_train.json"),
    ("8 GPUs, weak (64 pairs per GPU), final code", "r03s_n8_weak.json"),
    ("8 GPUs, strong (one batch of 64: 8 pairs per GPU), final code", "r03s_n8_strong.json"),
    ("8 GPUs, weak (64 pairs per GPU), earlier in the round", "r02s_n8_weak.json"),
    ("8 GPUs, strong (one batch of 64: 8 pairs per GPU), earlier in the round", "r02s_n8_strong.json"),
    ("8 GPUs, train (B=8 per GPU, graph)", "r02t_n8_train.json"),
]


def load(name):
    try:
        return json.loads(open(os.path.join(ROOT, "gpurun_out", name)).read().strip().splitlines()[-1])
    except Exception:
        return None


out = ["# Bench lines recorded in round 2 (B200; the full JSON lines live under gpurun_out/, which is not tracked)",
       "Single-GPU lines: run %s (final code of the round); 2- and 8-GPU inference lines: runs r03q / r03s (final code); the multi-GPU"
       % RUN, "training lines are earlier runs of the round (r02m / r02t).", ""]
for tag, f in RUNS:
    d = load(f)
    if not d:
        continue
    out.append("## %s  (%s)" % (tag, f))
    out.append("value %.1f %s, %.3f ms/step; e2e %.1f (%s); gpu_launches %s; clocks %s" % (
        d["value"], d["unit"], d["ms_per_step"], d["e2e"]["value"], d["e2e"].get("input", "host tensors"), d.get("gpu_launches"),
        d.get("clocks")))
    if "fp32_host_images" in d["e2e"]:
        out.append("e2e with fp32 host images: %.1f" % d["e2e"]["fp32_host_images"]["value"])
    if d.get("roofline"):
        out.append("roofline: %s" % {k: d["roofline"][k] for k in ("kernel", "bound", "achieved", "peak", "unit", "frac")})
    if d.get("kernels"):
        out.append("kernels (ms/step, frac of peak): %s" % {k: (v["ms_per_step"], v["frac"]) for k, v in d["kernels"].items()})
    if d.get("cpu_baseline"):
        c = d["cpu_baseline"]
        out.append("cpu_baseline: %.2f %s on %d cores (%s)" % (c["value"], c["unit"], c["cores"], c["kind"]))
    if d.get("torch_gpu_baseline"):
        out.append("torch_gpu_baseline: %s" % {k: v for k, v in d["torch_gpu_baseline"].items() if k != "sample"})
    if d.get("parity"):
        out.append("parity: %s" % d["parity"])
    if d.get("allreduce"):
        out.append("allreduce: %s" % d["allreduce"])
    if d.get("latency_ms_per_frame"):
        out.append("latency: %s" % d["latency_ms_per_frame"])
    out.append("")
open(os.path.join(ROOT, "profiles", TAG + "_bench_lines.md"), "w").write("\n".join(out))
print("\n".join(out))
