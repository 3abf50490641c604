"""Per-source-line summary of an ncu source page exported with
    ncu -i REP --page source --print-source cuda,sass --csv --kernel-name regex:NAME > out.csv
usage: python tools/ncu_lines.py out.csv [top]"""
import csv
import sys


def f(x):
    try:
        return float(x)
    except ValueError:
        return 0.0


rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out, fn, hdr = [], "", None
for r in rows:
    if r and r[0] == "File Path":
        fn = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) >= len(hdr) - 2 and r[0] not in ("", "Function Name"):
        out.append((fn, r))
ie, ss = hdr.index("Instructions Executed"), hdr.index("# Samples")
tie, ptie = hdr.index("Thread Instructions Executed"), hdr.index("Predicated-On Thread Instructions Executed")
tot = sum(f(r[ie]) for _, r in out)
tots = sum(f(r[ss]) for _, r in out)
print("warp instructions %.0f, samples %.0f" % (tot, tots))
for fn, r in sorted(out, key=lambda t: -f(t[1][ss]))[:top]:
    eff = f(r[ptie]) / max(f(r[ie]), 1)
    print("%6.2f%% inst %6.2f%% samp  lanes %4.1f | %-14s %4s | %s" % (100 * f(r[ie]) / tot, 100 * f(r[ss]) / tots, eff, fn, r[0],
                                                           r[1].strip()[:100]))
