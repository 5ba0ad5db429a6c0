"""Launch list (ncu --metrics gpu__time_duration.sum,... --csv) -> markdown: per-kernel-name totals and the launches in order.

  python tools/ncu_list.py TITLE gpurun_out/x.csv [LABEL2 gpurun_out/y.csv ...] > profiles/r02_x.md
"""
import csv
import sys


def load(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    h = rows[0]
    iK, iM, iV, iID = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
    k = {}
    for r in rows[1:]:
        k.setdefault(int(r[iID]), {"name": r[iK]})[r[iM]] = float(r[iV].replace(",", ""))
    return [k[i] for i in sorted(k)]


def short(n):
    n = n.split("(")[0].replace("void ", "").replace("ebk::<unnamed>::", "")
    return n[-70:]


def main():
    title, args = sys.argv[1], sys.argv[2:]
    print(f"# {title}\n\nTimes are cold-cache and serialised under ncu (`--clock-control none`); compare shares, not absolutes.\n")
    if len(args) == 1:
        args = ["step", args[0]]
    for label, path in zip(args[0::2], args[1::2]):
        ks = load(path)
        tot = sum(d["gpu__time_duration.sum"] for d in ks) / 1000
        agg = {}
        for d in ks:
            a = agg.setdefault(short(d["name"]), [0, 0.0])
            a[0] += 1
            a[1] += d["gpu__time_duration.sum"] / 1000
        print(f"## {label}: {len(ks)} launches, {tot:.0f} us\n\n| kernel | launches | us | share |\n|---|---:|---:|---:|")
        for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            print(f"| `{n}` | {c} | {t:.1f} | {100 * t / tot:.1f}% |")
        print()


main()
