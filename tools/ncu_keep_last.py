"""Keep the last N launches of an `ncu --csv` launch list (drops weight packing and warm-up in front of the timed forwards).

    python tools/ncu_keep_last.py all.csv N > last.csv
"""
import csv
import sys


def main(path, n):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    body = [r for r in rows[h + 1:] if len(r) > 5]
    csv.writer(sys.stdout).writerows(rows[:h + 1] + body[-n:])


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]))
