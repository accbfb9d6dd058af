"""Mnemonic counts per kernel of the built library (proof of what the SASS contains: UBLKCP / SYNCS = bulk-async copies +
mbarrier, FADD2 = packed fp32 adds, ATOMG / REDUX, fp64 ops ...):   python tools/sass_summary.py [lib.so] > profiles/..."""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "slambench_b200", "libkfb200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
WATCH = ["UBLKCP", "UTMALDG", "SYNCS", "FADD2", "FFMA2", "FMUL2", "MUFU", "REDUX", "ATOMG", "ATOMS", "RED", "LDG", "STG", "LDS", "STS",
         "SHFL", "BAR", "DFMA", "DADD", "DMUL", "CCTL", "NANOSLEEP"]
fun, counts, total = None, collections.OrderedDict(), {}
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fun = m.group(1); counts[fun] = collections.Counter(); total[fun] = 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
    if fun and m:
        total[fun] += 1
        op = m.group(1)
        for w in WATCH:
            if op == w or op.startswith(w + "."):
                counts[fun][w] += 1
for f in sorted(counts, key=lambda f: -total[f]):
    print(f"{f}: {total[f]} SASS instructions; " + ", ".join(f"{w} {counts[f][w]}" for w in WATCH if counts[f][w]))
