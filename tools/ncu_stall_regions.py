"""Per-region stall-sample breakdown of one kernel of an ncu report (SASS source page).

    python tools/ncu_stall_regions.py report.ncu-rep <kernel substring> [chunk size in instructions, default 100]

Prints, for every chunk of consecutive SASS instructions holding > 1 % of the warp-state samples, its share and top stall
reasons, then the fifteen most-sampled instructions.
"""
import collections
import csv
import io
import subprocess
import sys


def main(rep, kern, chunk=100):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    blocks, cur = [], None
    for line in out.splitlines():
        if line.startswith('"Kernel Name"'):
            cur = {"name": line, "lines": []}
            blocks.append(cur)
        elif cur is not None:
            cur["lines"].append(line)
    for b in blocks:
        if kern not in b["name"]:
            continue
        rows = list(csv.reader(io.StringIO("\n".join(b["lines"]))))
        hdr, data = rows[0], rows[1:]
        iS, iE, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
        sc = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        data = [r for r in data if len(r) > iSm and r[iSm].isdigit()]
        tot = sum(int(r[iSm]) for r in data)
        print(f"== {b['name'][:110]}  samples {tot}, instructions executed {sum(int(r[iE]) for r in data)}")
        mix = collections.Counter()
        for r in data:
            for c in sc:
                if r[c].isdigit():
                    mix[hdr[c][6:]] += int(r[c])
        print("   overall:", " ".join(f"{k}:{100 * v / tot:.0f}%" for k, v in mix.most_common(8)))
        for i in range(0, len(data), chunk):
            seg = data[i:i + chunk]
            sm = sum(int(r[iSm]) for r in seg)
            if sm > tot * 0.01:
                m = collections.Counter()
                for r in seg:
                    for c in sc:
                        if r[c].isdigit():
                            m[hdr[c][6:]] += int(r[c])
                print(f"   [{i:5d}] {100 * sm / tot:5.1f}%  " + " ".join(f"{k}:{v}" for k, v in m.most_common(3)) + f"   | {seg[0][iS].strip()[:50]}")
        idx = sorted(range(len(data)), key=lambda i: -int(data[i][iSm]))[:15]
        for i in sorted(idx):
            r = data[i]
            m = sorted(((int(r[c]), hdr[c][6:]) for c in sc if r[c].isdigit() and int(r[c])), reverse=True)[:2]
            print(f"   #{i:5d} smp {r[iSm]:>5s} exec {r[iE]:>7s}  {r[iS].strip()[:70]:70s} {m}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 100)
