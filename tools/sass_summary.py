#!/usr/bin/env python
"""SASS opcode summary of libspdm.so (runs on the CPU box: `python tools/sass_summary.py > profiles/r02_sass_opcodes.md`).

Counts, per kernel, the mnemonics that prove the Blackwell-native instructions (B200_PROFILING.md): UTCHMMA / UTCQMMA (tcgen05.mma),
UTMALDG / UTMASTG (TMA load / store), LDTM / STTM (tcgen05.ld / st), UTCBAR (tcgen05.commit), SYNCS (mbarrier), plus registers
from the ELF resource usage."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "state_policy_diffusionmodel_b200", "libspdm.so")
OPS = ("UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "STTM", "UTCBAR", "UTCATOM", "SYNCS", "HMMA", "MUFU", "UCGABAR")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    regs = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
        m = re.search(r"REG:(\d+).*SHARED:(\d+)", line)
        if m and cur:
            regs[cur] = (int(m.group(1)), int(m.group(2)))
    counts = collections.OrderedDict()
    name = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            counts[name] = collections.Counter()
            continue
        if name is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1).split(".")[0]
            counts[name]["_total"] += 1
            for o in OPS:
                if op == o or op.startswith(o + "_"):
                    counts[name][o] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
    total = collections.Counter()
    rows = []
    for (mangled, c), nice in zip(counts.items(), demangle):
        nice = nice.replace("(anonymous namespace)::", "")
        nice = re.sub(r"^void ", "", nice)
        nice = re.sub(r"\(.*", "", nice)
        for k, v in c.items():
            total[k] += v
        if any(c[o] for o in OPS if o not in ("MUFU", "SYNCS")):
            r, sm = regs.get(mangled, (0, 0))
            rows.append((nice, c, r, sm))
    print("# SASS opcode summary of libspdm.so (sm_100a), `cuobjdump -sass` of the built library\n")
    print("%d kernels in the library; the table lists those with tensor-core / TMA / TMEM instructions.\n" % len(counts))
    print("| kernel | instr | regs | static smem | " + " | ".join(o for o in OPS) + " |")
    print("|---|---|---|---|" + "---|" * len(OPS))
    for nice, c, r, sm in sorted(rows, key=lambda x: -x[1]["UTCHMMA"]):
        print("| `%s` | %d | %d | %d | " % (nice[:90], c["_total"], r, sm) + " | ".join(str(c[o]) if c[o] else "" for o in OPS) + " |")
    print("\nTotals over the whole library: " + ", ".join("%s %d" % (o, total[o]) for o in OPS) + ".")
    print("\n`UTCHMMA` = tcgen05.mma (kind::f16), `UTMALDG` = cp.async.bulk.tensor (TMA load), `LDTM` / `STTM` = tcgen05.ld / st, "
          "`UTCBAR` = tcgen05.commit, `UCGABAR` = barrier.cluster, `SYNCS` = mbarrier operations, `HMMA` = mma.sync (training attention core).")


if __name__ == "__main__":
    sys.exit(main())
