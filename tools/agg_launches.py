"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel name.
Usage: python tools/agg_launches.py file.csv [marker-kernel-substring]  (aggregates between the last two marker launches)"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    marker = sys.argv[2] if len(sys.argv) > 2 else "adam_kernel"
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    names = [r["Kernel Name"] for r in rows]
    idx = [i for i, n in enumerate(names) if marker in n]
    start, end = (idx[-2] + 1, idx[-1] + 1) if len(idx) >= 2 else (0, len(rows))
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for r in rows[start:end]:
        n = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("<unnamed>::", "")
        v = float(r["Metric Value"])
        u = r["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        agg[n][0] += 1
        agg[n][1] += v
        tot += v
    print("launches %d, total %.1f us" % (end - start, tot))
    for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
        print("%9.1f us %5.1f%% %4d  %s" % (v, 100 * v / tot, c, n[:100]))


if __name__ == "__main__":
    main()
