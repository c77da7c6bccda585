"""Per-kernel SASS opcode histogram of libsmx.so (the Blackwell-specific mnemonics the profiling recipe names).

    python tools/sass_histogram.py > profiles/r02_sass_opcodes.txt        # runs without a GPU (cuobjdump)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "summarymixing_b200", "libsmx.so")
OPS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "LDGSTS", "MUFU.TANH", "HMMA", "SYNCS", "UTCBAR"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            per[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            per[cur]["_total"] += 1
            for o in OPS:
                if op == o or op.startswith(o + "."):
                    per[cur][o] += 1
            if op.startswith("UTCHMMA") and ".2CTA" in op:
                per[cur]["UTCHMMA.2CTA"] += 1
    cols = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "LDGSTS", "MUFU.TANH", "HMMA"]
    print(f"# SASS opcode counts per kernel of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass); UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st,")
    print("# UTMALDG = cp.async.bulk.tensor (tensor-map TMA), UBLKCP = cp.async.bulk, LDGSTS = cp.async, HMMA = legacy mma.sync (must be 0)")
    print(f"{'kernel':72s} " + " ".join(f"{c:>12s}" for c in cols) + f" {'instructions':>12s}")
    tot = collections.Counter()
    for k, c in per.items():
        if not any(c[o] for o in cols):
            continue
        print(f"{k[:72]:72s} " + " ".join(f"{c[o]:12d}" for o in cols) + f" {c['_total']:12d}")
        tot.update(c)
    print(f"{'TOTAL (kernels listed)':72s} " + " ".join(f"{tot[o]:12d}" for o in cols) + f" {tot['_total']:12d}")


if __name__ == "__main__":
    main()
