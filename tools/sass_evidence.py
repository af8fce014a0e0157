#!/usr/bin/env python
"""Static SASS evidence for the built library (no GPU needed): per kernel the instruction count and
the counts of the mnemonics DESIGN.md refers to. Writes profiles/r02_sass_evidence.txt (round 1: r01_sass_evidence.txt).
    python tools/sass_evidence.py [out]"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "opensot_b200", "lib", "libopensot_b200.so")


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_evidence.txt")
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), [])
        elif cur is not None and re.search(r"/\*[0-9a-f]{4,}\*/\s+\S", line):
            cur.append(line)
    pat = [("UBLKCP", r"\bUBLKCP"), ("SYNCS", r"\bSYNCS\."), ("REDUX", r"\bC?REDUX"), ("LDS.128", r"\bLDS\.128"),
           ("DFMA", r"\bDFMA"), ("FFMA", r"\bFFMA"), ("SHFL", r"\bSHFL"), ("MUFU.RCP64H", r"MUFU\.RCP64H"),
           ("MUFU.RSQ64H", r"MUFU\.RSQ64H"), ("LDL+STL", r"\b(LDL|STL)\b"), ("HMMA/UTCMMA", r"\b(HMMA|UTCMMA|UTCHMMA)")]
    rows = ["SASS evidence for opensot_b200/lib/libopensot_b200.so (cuobjdump -sass, sm_100a), tools/sass_evidence.py",
            "per kernel: instructions | " + " | ".join(p[0] for p in pat),
            "(UBLKCP = cp.async.bulk, the TMA bulk copy of the level blocks; SYNCS = mbarrier; (C)REDUX = redux.sync arg-max/arg-min; "
            "MUFU.*64H = seeds of the lean reciprocal / reciprocal square root; LDL+STL = local memory (spills); tensor pipe: none, DESIGN.md 4.2)"]
    for name in sorted(kernels):
        body = kernels[name]
        rows.append("%-62s %6d | " % (name, len(body)) + " | ".join(str(sum(1 for l in body if re.search(p[1], l))) for p in pat))
    with open(out, "w") as f:
        f.write("\n".join(rows) + "\n")
    print("\n".join(rows))


if __name__ == "__main__":
    main()
