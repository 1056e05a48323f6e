#!/usr/bin/env python
"""Aggregate the last training step of an `ncu --metrics gpu__time_duration.sum --csv` launch list by stream and kernel.
Usage: python tools/agg_train.py gpurun_out/r02_train_launches_raw.csv [top]"""
import collections
import csv
import io
import re
import sys


def main():
    txt = open(sys.argv[1]).read()
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 24
    lines = [l for l in txt.splitlines() if not l.startswith("==")]
    rows = list(csv.DictReader(io.StringIO("\n".join(lines))))
    names = [r["Kernel Name"] for r in rows]
    idx = [i for i, n in enumerate(names) if "adam_kernel" in n]
    s, e = idx[-2] + 1, idx[-1] + 1
    bys = collections.defaultdict(lambda: collections.defaultdict(lambda: [0, 0.0]))
    for r in rows[s:e]:
        n = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void <unnamed>::", "").replace("<unnamed>::", "")
        t = float(r["Metric Value"]) / 1000.0
        bys[r["Stream"]][n][0] += 1
        bys[r["Stream"]][n][1] += t
    for st, agg in bys.items():
        tot = sum(v[1] for v in agg.values())
        print("STREAM %s: %.1f us in %d launches" % (st, tot, sum(v[0] for v in agg.values())))
        for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
            print("   %9.1f us %5.1f%% %4d  %s" % (t, 100 * t / tot, c, n[:100]))


if __name__ == "__main__":
    main()
