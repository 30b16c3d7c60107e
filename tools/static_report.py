#!/usr/bin/env python
"""Static evidence of what the shipped library contains, from the .so alone (no GPU): per kernel the register / stack /
static shared-memory use (`cuobjdump --dump-resource-usage`) and the count of the SASS mnemonics that prove the
Blackwell data path (`cuobjdump -sass`; names per /opt/skills/guides/B200_PROFILING.md):
    UTCHMMA   tcgen05.mma            UTMALDG  TMA tensor load (cp.async.bulk.tensor)
    LDTM/STTM tcgen05.ld / .st       UTCBAR   tcgen05.commit -> mbarrier (.MULTICAST for CTA pairs)
    SYNCS     mbarrier ops           HMMA     mma.sync (the dh = 32 attention kernel only)
    REDUX     warp reductions        LDL/STL  local-memory traffic (spills) -- expected 0 in the hot loops
Writes a markdown table to stdout:  python tools/static_report.py > profiles/r02_static_sass.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "comorag_b200", "lib", "libcomorag_b200.so")
MNEMONICS = ("UTCHMMA", "UTMALDG", "LDTM", "STTM", "UTCBAR", "SYNCS", "HMMA", "REDUX", "LDL", "STL")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def short(name: str) -> str:
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = re.sub(r"^void ", "", name)
    return re.sub(r"\(.*$", "", name).replace("crag::", "")


def main() -> None:
    if not os.path.exists(LIB):
        sys.exit(f"{LIB} missing: python -m comorag_b200.build")
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", line)
        if m and cur:
            usage[cur] = tuple(int(x) for x in m.groups())
            cur = None
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts = collections.defaultdict(collections.Counter)
    variants = collections.defaultdict(set)
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
        if m:
            op = m.group(1)
            for mn in MNEMONICS:
                if op == mn or (mn in ("UTCHMMA", "UTMALDG", "UTCBAR") and op.startswith(mn)):
                    counts[cur][mn] += 1
                    if mn in ("UTCHMMA", "UTCBAR", "UTMALDG", "LDTM"):
                        variants[cur].add(op + m.group(2))
    names = demangle(sorted(set(usage) | set(counts)))
    print("# Static contents of `comorag_b200/lib/libcomorag_b200.so` (sm_100a), from `tools/static_report.py`\n")
    print("Per kernel: registers / stack bytes / static shared bytes (dynamic shared memory is set at launch) and the count "
          "of the SASS\nmnemonics that identify the data path.  `LDL` / `STL` = local-memory loads / stores (spills or "
          "runtime-indexed arrays).\n")
    print("| kernel | regs | stack | " + " | ".join(MNEMONICS) + " | tcgen05 / TMA forms |")
    print("|---|---|---|" + "---|" * len(MNEMONICS) + "---|")
    for mangled in sorted(names, key=lambda n: short(names[n])):
        reg, stack, shared, local = usage.get(mangled, (0, 0, 0, 0))
        c = counts.get(mangled, {})
        forms = ", ".join(sorted(variants.get(mangled, ())))
        print(f"| `{short(names[mangled])}` | {reg} | {stack} | " + " | ".join(str(c.get(m, 0)) for m in MNEMONICS) + f" | {forms} |")
    total = collections.Counter()
    for c in counts.values():
        total.update(c)
    print("\nTotals: " + ", ".join(f"{m} {total[m]}" for m in MNEMONICS))


if __name__ == "__main__":
    main()
