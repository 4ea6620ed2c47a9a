"""Summarise an ncu report of a warp-specialised kernel: barrier-wait / TMA / MMA / TMEM sites with executed counts and
stall samples (development tool).  usage: python scripts/ncu_sites.py report.ncu-rep"""
import csv
import subprocess
import sys

rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr, data = rows[1], rows[2:]
iS, iE, iSrc = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
print("total samples", sum(int(r[iS]) for r in data), "warp-instructions", sum(int(r[iE]) for r in data))
keys = ("SYNCS.PHASECHK", "UTMALDG", "UTCHMMA", "UTCBAR", "LDTM", "BAR.SYNC", "SYNCS.ARRIVE")
for i, r in enumerate(data):
    if any(k in r[iSrc] for k in keys) and int(r[iE]) > 0:
        print("%5d samples=%6s exec=%9s  %s" % (i, r[iS], r[iE], r[iSrc].strip()[:90]))
top = sorted(range(len(data)), key=lambda i: -int(data[i][iS]))[:12]
print("-- top stall sites")
for i in top:
    print("%5d samples=%6s exec=%9s  %s" % (i, data[i][iS], data[i][iE], data[i][iSrc].strip()[:90]))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum", "l1tex__t_bytes.sum"]
for h, u, v in zip(rr[0], rr[1], rr[2]):
    if h in want:
        print(h, u, v)
