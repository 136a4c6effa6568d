"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (count, total,
mean, share).  Usage: python tools/summarize_launches.py gpurun_out/launches.csv [skip_first_n]"""
import csv
import sys
from collections import OrderedDict


def main(path, skip=0):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    k, v = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = OrderedDict()
    for r in rows[skip:]:
        name = r[k].split("(")[0].replace("void ", "")
        ns = float(r[v].replace(",", ""))
        c, t = agg.get(name, (0, 0.0))
        agg[name] = (c + 1, t + ns)
    total = sum(t for _, t in agg.values())
    print(f"| kernel | launches | total ms | mean ms | share |\n|---|---|---|---|---|")
    for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name[:90]}` | {c} | {t/1e6:.3f} | {t/c/1e6:.4f} | {t/total:.3f} |")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
