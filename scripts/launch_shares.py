"""Summarise an ncu launch list (gpu__time_duration per launch) of `bench.py`: kernel shares of the timed resident
step(s), captured with `CCVPE_NCU_RANGE=1 ncu --profile-from-start off ...` (scripts/gpu_round.sh).  Profiles tool."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
launches = []
for r in data:
    if len(r) <= mv:
        continue
    try:
        launches.append((r[kn], float(r[mv].replace(",", "")) / 1e3))   # us
    except ValueError:
        pass
# the list is captured with `ncu --profile-from-start off` around bench.py's timed resident step(s)
# (CCVPE_NCU_RANGE=1), so every launch in it belongs to the timed region
step = launches
agg = collections.defaultdict(lambda: [0, 0.0])
for n, us in step:
    key = n.split("(")[0].replace("void ", "")[:80]
    agg[key][0] += 1
    agg[key][1] += us
tot = sum(v[1] for v in agg.values())
ours = sum(v[1] for k, v in agg.items() if "ccvpe::" in k)
print("one resident step: %d launches, %.1f us serialized (cold-cache) kernel time; ccvpe kernels %.1f us (%.1f %%)" %
      (len(step), tot, ours, 100 * ours / tot))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print("%6.2f%% %5d %10.1f us  %s" % (100 * v[1] / tot, v[0], v[1], k))
