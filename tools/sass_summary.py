"""SASS evidence for the shipped library (read on the CPU box): per kernel, how many instructions of the kinds DESIGN.md
talks about the cubin contains -- TMA (UTMALDG / UTMASTG / UBLKCP), mbarrier waits (SYNCS), 256-bit global loads/stores,
reductions / atomics (RED / ATOMG / ATOMS), warp primitives (SHFL / MATCH / VOTE), FP64 arithmetic, and the division
slow path (CALL) that div_exact removed from the particle kernels.

usage: python tools/sass_summary.py [lib.so] > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "spacecharge.jl_b200", "lib", "libspacecharge_b200.so")
COLS = ["UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "LDG.256", "STG.256", "RED", "ATOMG", "ATOMS", "SHFL", "MATCH", "DFMA+DMUL+DADD",
        "MUFU", "CALL", "LDL+STL", "total"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    counts = collections.OrderedDict()
    cur = None
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", line)
        if not m or cur is None:
            continue
        op = m.group(1)
        c = counts[cur]
        c["total"] += 1
        base = op.split(".")[0]
        if base in ("UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "RED", "ATOMG", "ATOMS", "SHFL", "MATCH", "MUFU", "CALL"):
            c[base] += 1
        if base == "REDG":
            c["RED"] += 1
        if base == "LDG" and ".256" in op:
            c["LDG.256"] += 1
        if base == "STG" and ".256" in op:
            c["STG.256"] += 1
        if base in ("DFMA", "DMUL", "DADD"):
            c["DFMA+DMUL+DADD"] += 1
        if base in ("LDL", "STL"):
            c["LDL+STL"] += 1
    names = demangle(list(counts))
    print("cuobjdump -sass %s" % os.path.relpath(LIB, ROOT))
    print("cubin architectures: %s" % ", ".join(arch))
    tot = collections.Counter()
    for c in counts.values():
        tot.update(c)
    print("whole library: " + ", ".join("%s %d" % (k, tot[k]) for k in COLS))
    print()
    print("%-100s " % "kernel" + " ".join("%8s" % k[:8] for k in COLS))
    for k, c in counts.items():
        short = re.sub(r"\(.*", "", names[k]).replace("scb::", "").replace("void ", "")
        print("%-100s " % short[:100] + " ".join("%8d" % c[x] for x in COLS))


if __name__ == "__main__":
    main()
