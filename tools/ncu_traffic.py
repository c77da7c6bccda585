"""profiles/r02_ksm_traffic.json from the ncu csv of tools/ncu_cell_capture.py (metrics dram__bytes_read.sum,
dram__bytes_write.sum, gpu__time_duration.sum; --cache-control none --replay-mode application).

    python tools/ncu_traffic.py gpurun_out/ksm_traffic.csv profiles/r02_ksm_traffic.json
"""
import csv
import json
import sys


def main(src, dst):
    rows = []
    with open(src) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        rows.append(r)
    kern = {}
    order = []
    for r in rows:
        key = (r["ID"], r["Kernel Name"])
        if key not in kern:
            kern[key] = {}
            order.append(key)
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6}.get(unit, 1)
        kern[key][r["Metric Name"]] = v * scale
    launches = [{"id": k[0], "kernel": k[1][:80], **kern[k]} for k in order]
    cell = [l for l in launches if "cell" in l["kernel"] and "smx" in l["kernel"]]
    if not cell:
        raise SystemExit("no smx cell kernel in the capture")
    i_last = max(launches.index(c) for c in cell)
    evict = launches[i_last + 1] if i_last + 1 < len(launches) else None
    rd = sum(c.get("dram__bytes_read.sum", 0.0) for c in cell)
    wr = sum(c.get("dram__bytes_write.sum", 0.0) for c in cell)
    deferred = evict.get("dram__bytes_write.sum", 0.0) if evict else 0.0
    out = {"dram_bytes_per_call": rd + wr + deferred, "read_in_kernel": rd, "write_in_kernel": wr,
           "write_deferred_seen_in_evict_pass": deferred,
           "kernels": [c["kernel"] for c in cell], "kernel_ns": [c.get("gpu__time_duration.sum") for c in cell],
           "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --replay-mode application of "
                     "tools/ncu_cell_capture.py: L2 flushed by a 512 MB read before the call (cold x, cold weights); the output's "
                     "dirty lines still in L2 when the kernel ends are counted from the dram write bytes of a trailing 512 MB pure-read pass",
           "launches": launches}
    with open(dst, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps({k: out[k] for k in ("dram_bytes_per_call", "read_in_kernel", "write_in_kernel", "write_deferred_seen_in_evict_pass")}))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
