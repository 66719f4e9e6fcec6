"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total time and share."""
import collections
import csv
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path, errors="replace")))
    hdr = None
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        if len(r) > 5 and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            if d.get("Metric Name") == "gpu__time_duration.sum":
                v = float(d["Metric Value"].replace(",", ""))
                v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}[d["Metric Unit"]]
                k = d["Kernel Name"].split("(")[0][:90]
                agg[k][0] += 1
                agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {tot / 1e6:.3f} ms of kernel time (cold-cache, serialised)")
    print(f"{'total ms':>10} {'launches':>8} {'mean us':>9} {'share':>7}  kernel")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
        print(f"{t / 1e6:10.3f} {n:8d} {t / n / 1e3:9.2f} {100 * t / tot:6.1f}%  {k}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
