"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel share of device time."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]
    kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        v = float(r[mv].replace(",", ""))
        v = v / 1e3 if r[mu] in ("nsecond", "ns") else (v * 1e3 if r[mu] in ("msecond", "ms") else v)
        name = r[kn].split("(rpt::")[0].replace("void rpt::", "").replace("rpt::", "")
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:44s} launches={v[0]:4d} total_us={v[1]:10.1f} share={100 * v[1] / tot:5.1f}% avg_us={v[1] / v[0]:8.1f}")


if __name__ == "__main__":
    main(sys.argv[1])
