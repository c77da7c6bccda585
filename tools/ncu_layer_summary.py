"""Per-kernel summary (one CSV row per profiled launch) of an `ncu --set full` report: duration, DRAM bytes, pipe utilisation,
registers, shared memory, top stall reasons.

    python tools/ncu_layer_summary.py report.ncu-rep "comment line" > profiles/r02_ncu_layer_summary.csv
"""
import csv
import io
import subprocess
import sys

COLS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__cycles_active.avg", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
]


def main(rep, comment):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    stall = [(h, i) for h, i in idx.items() if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    w = csv.writer(sys.stdout)
    print(f'"# {comment}"')
    w.writerow(["Kernel Name"] + COLS + ["top stalls (share of the stalled-warp ratio sum)"])
    w.writerow([""] + [units[idx[c]] if c in idx else "" for c in COLS] + [""])
    for r in data:
        if len(r) <= idx["Kernel Name"]:
            continue
        name = r[idx["Kernel Name"]].replace("smx::", "").split("(CUtensorMap")[0].split("(smx::")[0].split("(const")[0]
        st = []
        for h, i in stall:
            try:
                st.append((float(r[i].replace(",", "")), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
        tot = sum(v for v, _ in st) or 1.0
        top = " ".join(f"{n}={100 * v / tot:.0f}%" for v, n in sorted(st, reverse=True)[:4])
        w.writerow([name] + [r[idx[c]].replace(",", "") if c in idx else "" for c in COLS] + [top])


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
