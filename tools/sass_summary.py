"""profiles/r02_sass_tma.txt: per kernel of the built library, how often the SASS mnemonics that show what it is made of
occur (cuobjdump -sass): TMA bulk copies and mbarrier operations, shared / global memory, FP64 arithmetic, warp-level
primitives, atomics, barriers.   python tools/sass_summary.py [out file]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "mot3d_b200", "libmot3d_b200.so")
out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_tma.txt")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEEP = ("UBLKCP", "SYNCS", "UTMALDG", "UTMASTG", "ATOMS", "ATOMG", "RED", "DADD", "DMUL", "DFMA", "MUFU", "BAR", "SHFL",
        "VOTE", "MATCH", "REDUX", "LDS", "STS", "LDG", "STG", "HMMA", "IMMA", "UTCHMMA", "UTCMMA")
fn, per = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        per[fn] = collections.Counter()
        continue
    if fn is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(2)
        base = op.split(".")[0]
        if base in KEEP:
            per[fn][op if base in ("UBLKCP", "SYNCS") else base] += 1
with open(out_path, "w") as f:
    f.write("cuobjdump -sass mot3d_b200/libmot3d_b200.so (sm_100a, built by __graft_entry__.build()): occurrences per kernel of the\n"
            "mnemonics that show what a kernel is made of. UBLKCP = cp.async.bulk (1-D TMA copy, global -> shared), SYNCS = mbarrier\n"
            "operations (arrive.expect_tx / try_wait); no UTMALDG: the staged slices are 1-D structure-of-arrays ranges, no tensor maps;\n"
            "no tensor-core instruction (HMMA / IMMA / UTC*MMA) anywhere: the path is not a contraction. D* = FP64 arithmetic of the\n"
            "reference's cost expressions; ATOMS / ATOMG / RED = shared / global atomics; MATCH / VOTE / REDUX / SHFL = warp primitives.\n\n")
    for name, c in per.items():
        if c:
            f.write("%s\n    %s\n" % (name[:160], "  ".join("%s=%d" % (k, v) for k, v in sorted(c.items()))))
print("wrote", out_path, len(per), "kernels")
