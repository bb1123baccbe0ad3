#!/usr/bin/env python
"""SASS evidence for the tensor / TMA paths: counts of the mnemonics that prove them, per kernel of libb200blas.so.
   python tools/sass_summary.py > profiles/rNN_sass_summary.txt
DMMA = FP64 tensor pipe; UTCHMMA/UTCQMMA... = tcgen05.mma; UTMALDG = TMA load; LDTM = tcgen05.ld; UTCBAR = tcgen05.commit;
SYNCS = mbarrier ops; UTMAPF = tensormap prefetch (B200_PROFILING.md)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "libgpublas_b200", "libb200blas.so")
KEYS = ["DMMA", "UTCHMMA", "UTCQMMA", "UTCIMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "SYNCS", "UTMAPF", "FFMA", "DFMA", "HMMA", "LDG", "STG", "LDS", "REDUX", "MULTIMEM"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1).split(".")[0]
            per[cur][op] += 1
    names = subprocess.run(["c++filt"], input="\n".join(per.keys()), capture_output=True, text=True).stdout.splitlines()
    total = collections.Counter()
    print("# cuobjdump -sass %s  (sm_100a)  -- instruction counts per kernel" % os.path.relpath(SO, ROOT))
    print("# %-110s %s" % ("kernel", " ".join("%s" % k for k in KEYS)))
    rows = []
    for (mangled, cnt), name in zip(per.items(), names):
        total.update(cnt)
        if not any(cnt[k] for k in ("DMMA", "UTCHMMA", "UTCQMMA", "UTMALDG", "LDTM", "UTCBAR")):
            continue
        rows.append((name, cnt))
    for name, cnt in sorted(rows):
        short = re.sub(r"\(.*", "", name)[:110]
        print("%-112s %s" % (short, " ".join("%s=%d" % (k, cnt[k]) for k in KEYS if cnt[k])))
    print("# TOTAL over all %d kernels: %s" % (len(per), " ".join("%s=%d" % (k, total[k]) for k in KEYS if total[k])))


if __name__ == "__main__":
    sys.exit(main())
