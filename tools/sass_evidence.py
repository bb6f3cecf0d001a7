#!/usr/bin/env python
"""Regenerates profiles/r02_sass_evidence.txt from the built library (no GPU needed):

    python tools/sass_evidence.py > profiles/r02_sass_evidence.txt

Per kernel of libgcxnorm.so: how many instructions, and how many of the mnemonics that show HOW it moves data and synchronises --
UBLKCP (cp.async.bulk, the TMA unit's bulk copy), UTMALDG (cp.async.bulk.tensor), UTC*MMA / LDTM / UTCBAR (tcgen05 tensor-core
products, TMEM loads, their commits), SYNCS (mbarrier), MEMBAR / ERRBAR / CCTL (fences of the peer exchange), the float64 pipe.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "gcmapexplorer_b200", "libgcxnorm.so")
WATCH = ["UBLKCP", "UTMALDG", "UTCIMMA", "UTCHMMA", "UTCQMMA", "LDTM", "UTCBAR", "SYNCS", "LDG.E", "STG.E", "LDGSTS", "DFMA", "DMUL", "DADD",
         "MUFU.RCP64H", "F2F.F64.F32", "F2F.F32.F64", "IMMA", "HMMA", "MEMBAR", "ERRBAR", "CCTL", "SHFL", "BAR.SYNC", "ATOM", "RED", "LDS", "STS",
         "LDL", "STL", "CALL"]
SHOW = ("UBLKCP", "UTMALDG", "UTCIMMA", "UTCHMMA", "LDTM", "UTCBAR", "SYNCS.ARRIVE", "MEMBAR", "ST.E.STRONG.SYS", "LD.E.STRONG.SYS", "CCTL")


def main():
    txt = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), [])
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?)\s*;", line)
        if m and cur is not None:
            cur.append(m.group(1))
    print("# SASS evidence (cuobjdump -sass gcmapexplorer_b200/libgcxnorm.so, sm_100a, CUDA 12.9) -- regenerate with tools/sass_evidence.py")
    print("# per kernel: instruction count and the mnemonics that show how data moves and how CTAs / GPUs synchronise, then the first")
    print("# occurrences of the telling ones.  UBLKCP = cp.async.bulk (TMA bulk copy), UTMALDG = cp.async.bulk.tensor, UTC*MMA / LDTM / UTCBAR =")
    print("# tcgen05.mma / tcgen05.ld / tcgen05.commit (only the correlation row's int8 products use the tensor cores: the normalization")
    print("# path is HBM-bound), SYNCS = mbarrier, MEMBAR / ERRBAR / CCTL = fences, DFMA / F2F.F64.F32 = float64 accumulation of float32 data.")
    total = collections.Counter()
    for name, ins in kernels.items():
        if not name.startswith("_ZN3gcx") and "gcx" not in name:
            continue
        cnt = collections.Counter()
        for i in ins:
            op = i.split()[1] if i.startswith("@") and len(i.split()) > 1 else i.split()[0]
            for w in WATCH:
                if op.startswith(w):
                    cnt[w] += 1
                    total[w] += 1
                    break
        print("\n## " + name)
        print("instructions: %d ; %s" % (len(ins), ", ".join("%s %d" % (w, cnt[w]) for w in WATCH if cnt[w])))
        shown = 0
        for i in ins:
            if any(s in i for s in SHOW) and shown < 8:
                print("    " + i)
                shown += 1
    print("\n## whole library")
    print(", ".join("%s %d" % (w, total[w]) for w in WATCH if total[w]))


if __name__ == "__main__":
    sys.exit(main())
