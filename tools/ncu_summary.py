"""Summarise an `ncu --set full` capture (`ncu -i X.ncu-rep --page raw --csv > X_raw.csv`) into the
handful of metrics DESIGN.md / bench.py cite.  Usage: python tools/ncu_summary.py X_raw.csv [title]"""
import csv
import sys

KEYS = [
    ("gpu__time_duration.sum", "kernel duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "registers/thread"),
    ("launch__occupancy_limit_registers", "occupancy limit (registers, blocks/SM)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy (% of 64 warps)"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "warp execution efficiency (active lanes / 32)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy (%)"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe (% of peak, inst issue)"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe cycles active (%)"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe (%)"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe (%)"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe (%)"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / cycle / SMSP"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput (% of peak)"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    title = sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]
    for k, vals in enumerate(rows[2:]):
        name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print(f"### {title} — launch {k}: `{name[:90]}`\n")
        print("| metric | value |\n|---|---|")
        for key, label in KEYS:
            if key in hdr:
                i = hdr.index(key)
                print(f"| {label} (`{key}`) | {vals[i]} {units[i]} |")
        stalls = []
        for i, h in enumerate(hdr):
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                try:
                    v = float(vals[i])
                except ValueError:
                    continue
                if v >= 0.05:
                    stalls.append((v, h.split("issue_stalled_")[1].split("_per_issue")[0]))
        stalls.sort(reverse=True)
        print("| stall reasons (warps stalled per issue) | " + ", ".join(f"{n} {v:.2f}" for v, n in stalls) + " |")
        print()


if __name__ == "__main__":
    main()
