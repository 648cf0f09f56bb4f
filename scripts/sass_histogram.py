"""SASS opcode histogram of libov2b200.so per kernel (cuobjdump -sass), written as markdown.
    python scripts/sass_histogram.py profiles/r2_sass_histogram.md
Rows of interest: UTMALDG (TMA tensor loads), SYNCS (mbarrier), DMMA (fp64 tensor core), RED / ATOMS / ATOMG, REDUX, IDP (dp4a/dp2a)."""
import collections
import re
import subprocess
import sys

dst = sys.argv[1]
lib = "ov2slam_b200/lib/libov2b200.so"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
fn, hist = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        fn = re.sub(r"\(anonymous namespace\)::", "", fn).split("(")[0]
        hist[fn] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and fn:
        hist[fn][m.group(1)] += 1
watch = ["UTMALDG", "UBLKCP", "SYNCS", "DMMA", "RED", "ATOMS", "ATOMG", "REDUX", "IDP", "PRMT", "LDG", "LDS", "STS", "STG", "DFMA", "DADD", "DMUL", "MUFU", "BAR", "SHFL"]
with open(dst, "w") as f:
    f.write("# SASS opcode histogram per kernel (cuobjdump -sass ov2slam_b200/lib/libov2b200.so, sm_100a)\n\n")
    f.write("Static instruction counts (not executed counts).  UTMALDG = `cp.async.bulk.tensor` TMA loads, SYNCS = mbarrier "
            "operations, DMMA = fp64 tensor-core MMA, RED = fire-and-forget global reductions.\n\n")
    f.write("| kernel | total | " + " | ".join(watch) + " |\n|---|---|" + "---|" * len(watch) + "\n")
    for fn, c in hist.items():
        f.write(f"| `{fn[:60]}` | {sum(c.values())} | " + " | ".join(str(c.get(w, 0)) for w in watch) + " |\n")
print(open(dst).read())
