#!/usr/bin/env python
"""Count the SASS mnemonics that prove which hardware paths each kernel uses (no GPU needed: reads the built objects).

  python tools/sass_summary.py > profiles/r01_sass_mnemonics.md

UTCHMMA / UTCQMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor load / store,
UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, UTCATOMSWS = TMEM alloc, REDG = global reduction atomics
(/opt/skills/guides/B200_PROFILING.md lists the mnemonics).
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "anoddpm_b200", "build")
KEYS = ("UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTCATOMSWS", "SYNCS", "REDG", "ATOMG",
        "LDS", "STS", "LDG", "STG", "LD.E", "ST.E", "MUFU", "HMMA", "FFMA", "DFMA", "SHFL")


def main():
    print("# SASS mnemonic counts per kernel (cuobjdump -sass of anoddpm_b200/build/*.o, sm_100a)\n")
    print("Static instruction counts, not dynamic ones: they show WHICH units a kernel drives (tcgen05 MMA issue, TMEM loads, TMA\n"
          "tensor loads / stores, mbarriers), not how often.\n")
    print("| object | kernel | " + " | ".join(KEYS) + " |\n|---|---|" + "---:|" * len(KEYS))
    for obj in sorted(os.listdir(BUILD)):
        if not obj.endswith(".o"):
            continue
        out = subprocess.run(["cuobjdump", "-sass", os.path.join(BUILD, obj)], capture_output=True, text=True).stdout
        cur, counts = None, collections.OrderedDict()
        for line in out.splitlines():
            m = re.match(r"\s*Function : (\S+)", line)
            if m:
                cur = m.group(1)
                counts[cur] = collections.Counter()
                continue
            m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if m and cur:
                op = m.group(1)
                for k in KEYS:
                    if op == k or op.startswith(k + ".") or (k in ("LD.E", "ST.E") and op.startswith(k)):
                        counts[cur][k] += 1
                        break
        for fn, c in counts.items():
            name = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(anonymous namespace\)::", "", name)
            name = re.sub(r"\(.*", "", name).replace("void ", "")
            if not any(c.values()):
                continue
            print(f"| {obj} | `{name}` | " + " | ".join(str(c.get(k, 0)) if c.get(k, 0) else "" for k in KEYS) + " |")


if __name__ == "__main__":
    sys.exit(main())
