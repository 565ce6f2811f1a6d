#!/usr/bin/env python
"""Per-kernel SASS evidence of the Blackwell-native paths in libmsml_b200.so (VERDICT r1 hygiene #14).

    python tools/sass_summary.py > profiles/r02_sass_summary.txt

Disassembles the in-tree library with `cuobjdump -sass` and counts, per kernel, the mnemonics that only sm_100a code
can contain: UTCHMMA (tcgen05.mma; .2CTA = cta_group::2), UTMALDG / UTMASTG (TMA tensor loads / stores), LDTM (tcgen05.ld),
UTCBAR (tcgen05.commit; .2CTA.MULTICAST = commit to both CTAs of a pair), UTCATOMSWS (TMEM allocation), plus the legacy
tensor path HMMA (must be zero) and the 128-bit global accesses of the streaming kernels.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "msml_b200", "libmsml_b200.so")
PATS = ["UTCHMMA.2CTA", "UTCHMMA", "UTMALDG.2D.2CTA", "UTMALDG", "UTMASTG", "LDTM", "UTCBAR.2CTA.MULTICAST", "UTCBAR", "UTCATOMSWS",
        "HMMA", "LDG.E.128", "STG.E.128", "LDGSTS", "RED.E", "MUFU.EX2", "SHFL"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def short(name):
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = re.sub(r"\bmsml::(tc::)?", "", name)
    name = re.sub(r"void ", "", name)
    return name[:150]


def main():
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
        m = re.search(r"^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        for p in PATS:
            if op.startswith(p):
                counts[cur][p] += 1
                break
    names = demangle(list(counts))
    print("# cuobjdump -sass msml_b200/libmsml_b200.so : Blackwell-native mnemonics per kernel (%d kernels)" % len(counts))
    print("# tcgen05.mma = UTCHMMA (cta_group::2 = .2CTA), TMA = UTMALDG/UTMASTG, tcgen05.ld = LDTM, tcgen05.commit = UTCBAR")
    tot = collections.Counter()
    rows = []
    for k, c in counts.items():
        tot.update(c)
        if any(c[p] for p in PATS[:9]):
            rows.append((short(names.get(k, k)), c))
    for name, c in sorted(rows, key=lambda r: r[0]):
        print("%s\n    %s" % (name, "  ".join("%s=%d" % (p, c[p]) for p in PATS if c[p])))
    print("\n# totals over all %d kernels" % len(counts))
    print("  ".join("%s=%d" % (p, tot[p]) for p in PATS))
    assert tot["HMMA"] == 0, "legacy mma.sync path found"
    return 0


if __name__ == "__main__":
    sys.exit(main())
