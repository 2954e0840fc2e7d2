"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv`
launch list of bench.py: DRAM bytes and device time of ONE factorization (the launches from the last
k_assemble_csc to the k_inertia that follows it) and of ONE solve sweep (k_perm_in .. k_perm_out),
per kernel.  usage: summarize_traffic.py traffic.csv"""
import collections
import csv
import sys


def main():
    rows = collections.OrderedDict()
    hdr = None
    for r in csv.reader(open(sys.argv[1], errors="replace")):
        if r and r[0] == "ID":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr):
            continue
        i = int(r[0])
        name = r[4].split("(")[0].replace("void ", "")
        d = rows.setdefault(i, {"name": name, "grid": r[8]})
        d[r[12]] = float(r[14].replace(",", ""))
    ids = sorted(rows)
    names = [rows[i]["name"] for i in ids]

    def last_span(first, last):
        e = max(k for k, n in enumerate(names) if n.endswith(last))
        b = max(k for k, n in enumerate(names[:e]) if n.endswith(first))
        return ids[b:e + 1]

    for title, first, last in (("factorization", "k_assemble_csc", "k_inertia"), ("solve sweep", "k_perm_in", "k_perm_out")):
        try:
            span = last_span(first, last)
        except ValueError:
            print(f"# {title}: not found")
            continue
        agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
        for i in span:
            d = rows[i]
            a = agg[d["name"]]
            a[0] += 1
            a[1] += d.get("dram__bytes_read.sum", 0.0)
            a[2] += d.get("dram__bytes_write.sum", 0.0)
            a[3] += d.get("gpu__time_duration.sum", 0.0)
        rd = sum(a[1] for a in agg.values()); wr = sum(a[2] for a in agg.values()); t = sum(a[3] for a in agg.values())
        print(f"# {title}: {len(span)} launches ({first} .. {last}), dram read {rd / 1e6:.1f} MB, write {wr / 1e6:.1f} MB "
              f"-> {(rd + wr) / 1e6:.1f} MB, cold serialised {t / 1e6:.3f} ms")
        print(f"  {'kernel':40s} {'launches':>8s} {'DRAM MB':>10s} {'us':>10s}")
        for k, a in sorted(agg.items(), key=lambda kv: -(kv[1][1] + kv[1][2])):
            print(f"  {k:40s} {a[0]:8d} {(a[1] + a[2]) / 1e6:10.1f} {a[3] / 1e3:10.1f}")


if __name__ == "__main__":
    main()
