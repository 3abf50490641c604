"""One line per bench JSON line (files given as arguments, else stdin): value, step, solve stage, k_ssp_solve, e2e."""
import fileinput
import json

for line in fileinput.input():
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    ssp = [k["ms"] for k in d.get("roofline_kernels", []) if k["kernel"].startswith("k_ssp_solve")]
    print("value %.1f M/s  step %.3f ms  solve %.3f ms  k_ssp %.3f ms  e2e %.3f ms" % (
        d["value"] / 1e6, d["ms_per_step"], d["stage_ms"]["solve"], ssp[0] if ssp else -1, d["e2e"]["ms_per_step"]))
