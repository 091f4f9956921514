"""Per-kernel SASS mnemonic counts of bcnf_b200/libbcnf_b200.so (cuobjdump -sass), plus ptxas resource usage.

    python profiles/sass_summary.py > profiles/r02_sass_summary.txt

What the mnemonics evidence (B200_PROFILING.md): UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM = tcgen05.ld,
UBLKCP = cp.async.bulk (TMA bulk copy), SYNCS = mbarrier ops, UCGABAR = cluster barrier, CCTL = cache control
(discard.global.L2 shows as CCTL.E.RML2), STG.E.ENL2.256 / LDG...256 = 32-byte sector stores / streaming loads,
FFMA2 = packed fp32x2 FMA, REDG/ATOMG = global reductions (log-det, ranks), LDL/STL = local memory (spills).
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "bcnf_b200", "libbcnf_b200.so")
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "UBLKCP", "SYNCS", "UCGABAR", "ELECT", "CCTL", "STG", "LDG", "LDS", "STS", "LDL", "STL",
        "FFMA2", "FFMA", "MUFU", "REDG", "ATOMG", "ATOMS", "MEMBAR", "HMMA"]


def demangle(name: str) -> str:
    try:
        return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().split("(")[0]
    except OSError:
        return name


def main() -> None:
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    cur, counts = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        counts[cur]["instructions"] += 1
        for k in KEYS:
            if op.startswith(k):
                counts[cur][k] += 1
                break
        if op.startswith("STG") and ".256" in op:
            counts[cur]["STG.256"] += 1
        if op.startswith("LDG") and ".256" in op:
            counts[cur]["LDG.256"] += 1
        if op.startswith("CCTL") and "RML2" in op:
            counts[cur]["CCTL.E.RML2"] += 1
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    fn = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            fn = m.group(1)
            continue
        if fn and "REG:" in line:
            usage[fn] = line.strip()
            fn = None
    print(f"# SASS summary of {os.path.relpath(LIB, ROOT)} (sm_100a), {len(counts)} kernels\n")
    for f, c in counts.items():
        print(demangle(f))
        print("   " + "  ".join(f"{k}={v}" for k, v in c.items()))
        if f in usage:
            print("   " + usage[f])
    sys.stdout.flush()


if __name__ == "__main__":
    main()
