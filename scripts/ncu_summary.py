"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: share of the step, count, mean time."""
import collections, csv, re, sys


def main(path, top=30):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg, n = collections.OrderedDict(), 0
    for row in r:
        if len(row) <= vi:
            continue
        v = float(row[vi].replace(",", ""))
        u = row[ui]
        us = v / 1000.0 if u.startswith("n") else (v if u.startswith("u") else v * 1000.0 if u.startswith("m") else v * 1e6)
        name = re.sub(r"\(.*", "", row[ki])
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
        n += 1
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: {n} launches, {tot / 1e3:.2f} ms of kernel time (cold-cache, serialised by ncu: compare SHARES)")
    print(f"{'share':>7} {'total ms':>10} {'launches':>9} {'mean us':>9}  kernel")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{a[1] / tot * 100:6.2f}% {a[1] / 1e3:10.2f} {a[0]:9d} {a[1] / a[0]:9.2f}  {k[:100]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
