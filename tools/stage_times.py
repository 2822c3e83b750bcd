"""Per-stage kernel times of the streaming integrator from an ncu launch list
(ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv python tools/prof_cfg.py ...).
usage: python tools/stage_times.py X.csv"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5 and r[0].isdigit()]
seq = [(r[4], float(r[-1].replace(",", ""))) for r in rows]
gens = [i for i, (k, _) in enumerate(seq) if "generate" in k]
start = gens[1] if len(gens) > 1 else 0          # skip the warm-up render call
tot = collections.defaultdict(float)


def short(k):
    m = re.search(r"k_stream_(\w+)(<[^>]*>)?", k)
    if not m:
        return k[:30]
    tp = m.group(2) or ""
    return m.group(1) + ("<any>" if m.group(1) == "trace" and tp.startswith("<(bool)1") else "")


for k, v in seq[start:]:
    tot[short(k)] += v
T = sum(tot.values())
print(f"total {T / 1e6:.3f} ms over {len(gens) - 1 if len(gens) > 1 else 1} wave(s) (ncu: serialized, cold)")
for k, v in sorted(tot.items(), key=lambda x: -x[1]):
    print(f"   {k:16s} {v / 1e6:8.3f} ms {100 * v / T:5.1f}%")
w = seq[start:gens[2]] if len(gens) > 2 else seq[start:]
print("   first wave:", " ".join(f"{short(k)}={v / 1e6:.2f}" for k, v in w[:24]))
