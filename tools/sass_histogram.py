#!/usr/bin/env python
"""Per-kernel SASS instruction histogram of the built library (cuobjdump -sass), so that the
Blackwell-native claim (UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA loads / stores, LDTM = tcgen05.ld,
UTCBAR = tcgen05.commit, SYNCS = mbarrier) can be checked without rebuilding.

    python tools/sass_histogram.py [lib.so] > profiles/r2_sass_histogram.txt
"""
from __future__ import annotations

import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
BLACKWELL = ("UTCHMMA", "UTCQMMA", "UTCOMMA", "UTMALDG", "UTMASTG", "UTMAPF", "UTMACMDFLUSH", "LDTM", "STTM", "UTCBAR",
             "UTCATOMSWS", "UTCCP", "SYNCS", "ELECT", "UCGABAR", "ACQBULK", "PREEXIT")
INSTR = re.compile(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P[0-9T]+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Za-z0-9_]+)*)")


def demangle(names):
    try:
        out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True, check=True).stdout.splitlines()
        return dict(zip(names, out))
    except Exception:
        return {n: n for n in names}


def main() -> int:
    lib = Path(sys.argv[1]) if len(sys.argv) > 1 else ROOT / "sr_object_detection_b200" / "libyolo2_b200.so"
    sass = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        if "Function :" in line:
            cur = line.split("Function :", 1)[1].strip()
            per[cur] = collections.Counter()
            continue
        m = INSTR.match(line)
        if m and cur is not None:
            full = m.group(1)
            base = full.split(".", 1)[0]
            # Blackwell-specific instructions keep their modifiers (2CTA, tile dims ...), the rest is grouped by opcode
            per[cur][full if base in BLACKWELL else base] += 1
    names = demangle(list(per))
    print(f"# SASS instruction histogram of {lib.name} ({len(per)} kernels), static instruction counts")
    print("# UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), UTMALDG/UTMASTG = TMA load/store, LDTM = tcgen05.ld,")
    print("# UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, UCGABAR = cluster barrier, ACQBULK/PREEXIT = programmatic dependent launch")
    tot_bw = collections.Counter()
    for fn, c in per.items():
        name = re.sub(r"\(.*", "", names[fn])
        total = sum(c.values())
        bw = {k: v for k, v in c.items() if k.split(".", 1)[0] in BLACKWELL}
        for k, v in bw.items():
            tot_bw[k.split(".", 1)[0] + (".2CTA" if ".2CTA" in k else "")] += v
        top = ", ".join(f"{k} {v}" for k, v in c.most_common(12) if k not in bw)
        print(f"\n{name}  [{total} instructions]")
        if bw:
            print("  blackwell: " + ", ".join(f"{k} {v}" for k, v in sorted(bw.items())))
        print("  top:       " + top)
    print("\n# library totals of the Blackwell-specific opcodes")
    for k, v in sorted(tot_bw.items()):
        print(f"#   {k:16s} {v}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
