"""Static evidence from the shipped library (no GPU needed):

  python profiles/tools/sass_evidence.py resources > profiles/r2_final_resources.txt
  python profiles/tools/sass_evidence.py sass      > profiles/r2_final_sass_blackwell.txt

resources: registers / stack (non-zero = local-memory spills or arrays) / static shared memory per kernel from
`cuobjdump --dump-resource-usage`; sass: counts of the Blackwell-specific mnemonics per kernel from `cuobjdump -sass`.
"""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "efficientdet_b200", "libeffdet_b200.so")
MNEMONICS = ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTMACCTL", "SYNCS", "FFMA2", "LDGSTS",
             "ELECT", "REDUX", "UCGABAR_ARV", "HFMA2")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), stdout=subprocess.PIPE, text=True).stdout.splitlines()
    short = []
    for n in out:
        n = re.sub(r"^void ", "", n).replace("effdet::", "")
        depth, cut = 0, len(n)
        for i, ch in enumerate(n):            # drop the parameter list, keep template arguments
            if ch == "<":
                depth += 1
            elif ch == ">":
                depth -= 1
            elif ch == "(" and depth == 0:
                cut = i
                break
        short.append(n[:cut])
    return short


def resources():
    txt = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], stdout=subprocess.PIPE, text=True).stdout
    rows = []
    for m in re.finditer(r"Function (\S+):\n\s+REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+) CONSTANT\[0\]:(\d+)", txt):
        rows.append((m.group(1), int(m.group(2)), int(m.group(3)), int(m.group(4)), int(m.group(5))))
    names = demangle([r[0] for r in rows])
    print("# cuobjdump --dump-resource-usage efficientdet_b200/libeffdet_b200.so (sm_100a); %d kernels" % len(rows))
    print("# STACK > 0 = local-memory frame (spills or indexed local arrays); SHARED = static shared memory (dynamic")
    print("# shared memory of the TMA / tcgen05 kernels is set at launch)")
    print("%-78s %5s %6s %7s" % ("kernel", "regs", "stack", "shared"))
    for n, r in sorted(zip(names, rows), key=lambda t: t[0]):
        print("%-78s %5d %6d %7d" % (n[:78], r[1], r[2], r[3]))
    spills = [n for n, r in zip(names, rows) if r[2] > 0]
    print("# kernels with a stack frame: %d %s" % (len(spills), sorted(set(spills))))


def sass():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True).stdout
    cur, counts, order = None, collections.defaultdict(collections.Counter), []
    for line in txt.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            cur = m.group(1)
            order.append(cur)
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            counts[cur]["_n"] += 1
            for k in MNEMONICS:
                if op == k or op.startswith(k):
                    counts[cur][k] += 1
    names = demangle(order)
    print("# cuobjdump -sass efficientdet_b200/libeffdet_b200.so: Blackwell-specific mnemonics per kernel")
    print("# UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UTMALDG / UTMASTG = TMA tensor load / store,")
    print("# UTMAPF = TMA L2 prefetch, SYNCS = mbarrier, ELECT = elect.sync, FFMA2 = packed fp32x2 FMA, REDUX = warp reduce,")
    print("# UCGABAR_ARV = cluster barrier, LDGSTS = cp.async")
    plain = []
    for f, n in sorted(zip(order, names), key=lambda t: t[1]):
        c = counts[f]
        tags = "  ".join("%s=%d" % (k, c[k]) for k in MNEMONICS if c[k])
        if tags:
            print("%-70s instr=%-6d %s" % (n[:70], c["_n"], tags))
        else:
            plain.append(n)
    print("# plain SIMT kernels (none of the above): %d" % len(plain))


if __name__ == "__main__":
    {"resources": resources, "sass": sass}[sys.argv[1]]()
