"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, mean, share.

    python tools/ncu_launch_summary.py launches.csv [last N launches only]"""
import collections
import csv
import sys


def main(path, last=0):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    h = rows[hdr]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    d = collections.defaultdict(list)
    body = [r for r in rows[hdr + 1:] if len(r) > vi]
    if last:  # only the last `last` launches (the timed steps: everything before is weight packing and warm-up)
        body = body[-last:]
    for r in body:
        if len(r) > vi:
            try:
                d[r[ki].split("(")[0][:70]].append(float(r[vi].replace(",", "")))
            except ValueError:
                pass
    tot = sum(sum(v) for v in d.values())
    print(f"{'total us':>10} {'n':>5} {'mean us':>9} {'share':>6}  kernel   [{path}: {sum(len(v) for v in d.values())} launches, {tot / 1e3:.1f} us]")
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        print(f"{sum(v) / 1e3:10.1f} {len(v):5d} {sum(v) / len(v) / 1e3:9.1f} {100 * sum(v) / tot:5.1f}%  {k}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
