"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total
device time and share.  usage: summarize_launches.py launches.csv [first_id last_id]"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    hi = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 60
    hdr, agg, n = None, collections.defaultdict(lambda: [0, 0.0, 0]), 0
    for r in csv.reader(open(path, errors="replace")):
        if r and r[0] == "ID":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr):
            continue
        i = int(r[0])
        if i < lo or i > hi:
            continue
        name = r[4].split("(")[0].replace("void ", "")
        t = float(r[-1].replace(",", ""))
        grid = int(r[8].strip("()").split(",")[0])
        agg[name][0] += 1
        agg[name][1] += t
        agg[name][2] += grid
        n += 1
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: launches {n}, total device time {tot / 1e6:.3f} ms (ncu: serialised, cold cache)")
    print(f"{'kernel':34s} {'launches':>8s} {'CTAs':>10s} {'ms':>10s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:34s} {v[0]:8d} {v[2]:10d} {v[1] / 1e6:10.3f} {100 * v[1] / tot:6.1f}%")


if __name__ == "__main__":
    main()
