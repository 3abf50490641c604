"""Top SASS instructions by stall samples of an ncu source page exported with
    ncu -i REP --page source --print-source cuda,sass --csv [--launch-skip i --launch-count 1] > out.csv
usage: python tools/ncu_sass.py out.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = None
cur_line = ""
out = []
for r in rows:
    if r and r[0] == "Line No":
        hdr = r
        continue
    if not hdr or len(r) < len(hdr) - 2:
        continue
    if r[0] not in ("", "Function Name"):
        cur_line = "%s: %s" % (r[0], r[1].strip()[:60])
    elif r[0] == "" and r[2].startswith("0x"):
        out.append((cur_line, r))
ss = hdr.index("# Samples")
stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(float(r[ss] or 0) for _, r in out)
print("samples", tot)
for line, r in sorted(out, key=lambda t: -float(t[1][ss] or 0))[:top]:
    st = sorted(((float(r[i] or 0), hdr[i][6:]) for i in stalls), reverse=True)[:2]
    print("%5.1f%% %-48s | %-40s | %s" % (100 * float(r[ss] or 0) / tot, r[3].strip()[:48], line[:40],
                                        " ".join("%s=%d" % (n, v) for v, n in st if v > 0)))
