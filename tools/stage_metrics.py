"""Per-stage table of the streaming integrator from an ncu metrics pass:
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file X.csv python tools/prof_cfg.py ...
(round 2: add lts__t_bytes.sum,lts__t_bytes.sum.pct_of_peak_sustained_elapsed,l1tex__t_bytes.sum,l1tex__t_bytes.sum.pct_of_peak_sustained_elapsed
to the metric list for the L2 / L1TEX columns)
usage: python tools/stage_metrics.py X.csv [hbm_peak_gbs] [rays_in_the_profiled_pass]"""
import collections
import csv
import re
import sys

peak = float(sys.argv[2]) if len(sys.argv) > 2 else 6459.3
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5 and r[0].isdigit()]
launch = collections.OrderedDict()
for r in rows:
    launch.setdefault(int(r[0]), {"name": r[4]})[r[-3]] = float(r[-1].replace(",", ""))
    launch[int(r[0])]["unit:" + r[-3]] = r[-2]


def short(k):
    m = re.search(r"k_stream_(\w+)(<[^>]*>)?", k)
    if not m:
        return None
    tp = m.group(2) or ""
    return m.group(1) + ("<any>" if m.group(1) == "trace" and tp.startswith("<(bool)1") else "")


ids = list(launch)
gens = [i for i in ids if short(launch[i]["name"]) == "generate"]
start = gens[1] if len(gens) > 1 else gens[0]
agg = collections.OrderedDict()
for i in ids:
    if i < start:
        continue
    d = launch[i]
    k = short(d["name"])
    if k is None:
        continue
    a = agg.setdefault(k, collections.defaultdict(float))
    t = d["gpu__time_duration.sum"]
    t_ms = t / 1e6 if d["unit:gpu__time_duration.sum"] in ("ns", "nsecond") else t / 1e3 if d["unit:gpu__time_duration.sum"] in ("us", "usecond") else t
    a["n"] += 1; a["ms"] += t_ms

    def to_bytes(key):
        u = d["unit:" + key]
        return d[key] * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    a["bytes"] += to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
    for key, nm in (("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma"), ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes"),
                    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue")):
        a[nm] += d[key] * t_ms
    a["regs"] = d["launch__registers_per_thread"]
    if "lts__t_bytes.sum" in d:
        a["l2"] += to_bytes("lts__t_bytes.sum"); a["l1"] += to_bytes("l1tex__t_bytes.sum")
        a["l2pct"] += d.get("lts__t_bytes.sum.pct_of_peak_sustained_elapsed", 0.0) * t_ms
        a["l1pct"] += d.get("l1tex__t_bytes.sum.pct_of_peak_sustained_elapsed", 0.0) * t_ms
T = sum(a["ms"] for a in agg.values())
print(f"| stage kernel | launches | time ms (share) | regs | occupancy % of 64 warps | lanes / 32 | issue slots % | FMA pipe % | DRAM GB/s (of {peak:.0f}) |")
print("|---|---|---|---|---|---|---|---|---|")
for k, a in agg.items():
    print(f"| {k} | {int(a['n'])} | {a['ms']:.2f} ({100 * a['ms'] / T:.0f} %) | {int(a['regs'])} | {a['occ'] / a['ms']:.0f} | {a['lanes'] / a['ms']:.1f} | {a['issue'] / a['ms']:.0f} | "
          f"{a['fma'] / a['ms']:.0f} | {a['bytes'] / a['ms'] / 1e6:.0f} ({100 * a['bytes'] / a['ms'] / 1e6 / peak:.0f} %) |")
if any(a["l2"] for a in agg.values()):
    rays = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
    print("\n| stage kernel | L2 traffic GB/s (% of the L2's peak, ncu) | L1TEX traffic GB/s (% of peak, ncu) | L2 bytes | L1TEX bytes |")
    print("|---|---|---|---|---|")
    for k, a in agg.items():
        print(f"| {k} | {a['l2'] / a['ms'] / 1e6:.0f} ({a['l2pct'] / a['ms']:.0f} %) | {a['l1'] / a['ms'] / 1e6:.0f} ({a['l1pct'] / a['ms']:.0f} %) | {a['l2'] / 1e6:.0f} MB | {a['l1'] / 1e6:.0f} MB |")
    if rays:
        tr = [a for k, a in agg.items() if k.startswith("trace") or k in ("closest", "shadow")]
        print(f"\nintersection kernels (closest, shadow, trace): {sum(a['l1'] for a in tr) / rays:.0f} B of L1TEX and {sum(a['l2'] for a in tr) / rays:.0f} B of L2 traffic per ray "
              f"({rays / 1e6:.1f} M rays); whole wave {rays / (T * 1e-3) / 1e9:.2f} Grays/s under ncu's serialised, cold-cache launches")
tot_bytes = sum(a["bytes"] for a in agg.values())
print(f"\ntotal {T:.2f} ms, {tot_bytes / 1e9:.2f} GB of DRAM traffic = {tot_bytes / T / 1e6:.0f} GB/s average")
