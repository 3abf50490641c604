"""Prints the latency section of a bench line (or the JSON lines of tools/window_latency.py) as a table.
    python tools/latency_table.py FILE"""
import json
import sys

rows = []
for line in open(sys.argv[1]):
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    rows += d["latency"] if "latency" in d else [d]
for r in rows:
    gpu = ", ".join("%s %.3f / %.3f" % (k, v["p50"], v["p99"]) for k, v in r["gpu_ms"].items())
    cpu = r.get("cpu_ms", {}).get("p50")
    txt = r.get("cpu_text_ms", {}).get("p50")
    print("%-30s %7d dets  gpu ms p50/p99: %-52s cpu %s  text %s  agree %s" % (
        r["case"], r.get("detections", 0), gpu, "%.2f" % cpu if cpu else "-", "%.2f" % txt if txt else "-",
        r.get("totals_agree")))
