#!/usr/bin/env python
"""Per-kernel SASS opcode census of libdalm_b200.so (cuobjdump -sass), written to profiles/rNN_sass_opcodes.txt.

    python tools/sass_report.py [out_path]

Shows which kernels run on the 5th-gen tensor cores (UTCHMMA = tcgen05.mma, .2CTA = cta_group::2; LDTM/STTM = tcgen05.ld/st
to TMEM; UTMALDG/UTMASTG = TMA tensor loads/stores) and which still use warp-level HMMA (mma.sync)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "dalm_b200", "csrc", "libdalm_b200.so")
OPS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "HMMA", "LDGSTS", "LDSM", "FFMA2", "MUFU.EX2"]


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_opcodes.txt")
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"^\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        c = counts[cur]
        c["_total"] += 1
        for o in OPS:
            if o == "UTCHMMA.2CTA":
                if op.startswith("UTCHMMA") and ".2CTA" in op:
                    c[o] += 1
            elif op == o or op.startswith(o + "."):
                c[o] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
    rows = []
    for (name, c), dn in zip(counts.items(), demangle):
        short = re.sub(r"\(.*", "", dn).replace("dalm::", "")
        rows.append((short, c))
    rows.sort(key=lambda r: (-r[1]["UTCHMMA"], -r[1]["HMMA"], r[0]))
    tot = collections.Counter()
    with open(out, "w") as f:
        f.write(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} — instruction counts per kernel (static SASS, sm_100a)\n")
        f.write("# UTCHMMA = tcgen05.mma (…2CTA = cta_group::2), LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA load/store, HMMA = mma.sync\n")
        f.write(f"{'kernel':<58}" + "".join(f"{o:>13}" for o in OPS) + f"{'total':>9}\n")
        for short, c in rows:
            f.write(f"{short[:57]:<58}" + "".join(f"{c[o]:>13}" for o in OPS) + f"{c['_total']:>9}\n")
            tot.update(c)
        f.write(f"{'ALL KERNELS':<58}" + "".join(f"{tot[o]:>13}" for o in OPS) + f"{tot['_total']:>9}\n")
    print(out)
    print({o: tot[o] for o in OPS})


if __name__ == "__main__":
    main()
