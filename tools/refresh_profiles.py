"""Turns the files of an evidence run (gpurun_out/<tag>_bench.json, _bench_reference_arm.json, _launches_bench.csv,
_full.ncu-rep) into the committed summaries under profiles/: the bench lines, the ncu launch list, the raw metrics of
the full capture and profiles/ncu_traffic.json (DRAM bytes per kernel, read by bench.py).
    python tools/refresh_profiles.py r02c [round-prefix, default r02]"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
rnd = sys.argv[2] if len(sys.argv) > 2 else "r02"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
shutil.copy(os.path.join(G, tag + "_bench.json"), os.path.join(P, rnd + "_bench.json"))
shutil.copy(os.path.join(G, tag + "_bench_reference_arm.json"), os.path.join(P, rnd + "_bench_reference_arm.json"))
shutil.copy(os.path.join(G, tag + "_launches_bench.csv"), os.path.join(P, rnd + "_launches_bench.csv"))
metrics = ("gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,"
           "sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,"
           "launch__grid_size,launch__block_size,dram__throughput.avg.pct_of_peak_sustained_elapsed,"
           "sm__cycles_active.avg,sm__cycles_elapsed.avg")
out = subprocess.run(["ncu", "-i", os.path.join(G, tag + "_full.ncu-rep"), "--page", "raw", "--csv", "--metrics", metrics],
                     capture_output=True, text=True).stdout
open(os.path.join(P, rnd + "_ncu_full_metrics.csv"), "w").write(out)
rows = list(csv.reader(out.splitlines()))
h = rows[0]
per = collections.OrderedDict()
for r in rows[2:]:
    d = dict(zip(h, r))
    name = d["Kernel Name"].split("(")[0].replace("void ", "")
    per.setdefault(name, []).append({
        "dram": (float(d["dram__bytes_read.sum"]) + float(d["dram__bytes_write.sum"])) * 1e6,
        "ms": float(d["gpu__time_duration.sum"]), "inst": float(d["smsp__inst_executed.sum"]),
        "issue": float(d["smsp__issue_active.avg.pct_of_peak_sustained_active"]),
        "grid": d["launch__grid_size"], "block": d["launch__block_size"]})
print("%-34s %8s %8s %9s %7s %s" % ("kernel", "ms", "DRAM MB", "instr M", "issue%", "grid x block"))
tr = {}
for k, v in per.items():
    m = {f: sum(x[f] for x in v) / len(v) for f in ("dram", "ms", "inst", "issue")}
    tr[k] = m["dram"]
    print("%-34s %8.3f %8.0f %9.0f %7.1f %s x %s" % (k, m["ms"], m["dram"] / 1e6, m["inst"] / 1e6, m["issue"],
                                                  v[0]["grid"], v[0]["block"]))


def pick(prefix):
    return sum(v for k, v in tr.items() if k.startswith(prefix))


ssp = pick("k_ssp_solve")
traffic = {
    "comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full capture profiles/%s_ncu_full_"
               "metrics.csv (tools/profile_solve.py 1200 = bench.py's default configuration: 60 sequences/GPU = 1200 graphs, "
               "6.0 M detections, 47.9 M arcs per step). k_ssp_solve = sum over its band launches; solve_stage = k_first_tree"
               " + k_components_tree + k_single_track + k_job_order + k_ssp_solve; edge_build = k_frame_sort + "
               "k_edge_build_fused" % rnd,
    "seqs": 60, "k_ssp_solve": ssp,
    "solve_stage": pick("k_first_tree") + pick("k_components") + pick("k_single_track") + pick("k_job_order") + ssp,
    "edge_build": pick("k_frame_sort") + pick("k_edge_build_fused"),
    "per_kernel": {"k_frame_sort": pick("k_frame_sort"), "k_edge_build_fused": pick("k_edge_build_fused"),
                   "k_first_tree": pick("k_first_tree"), "k_components_tree": pick("k_components"),
                   "k_single_track": pick("k_single_track") + pick("k_job_order"), "k_ssp_solve": ssp,
                   "k_extract_tracks": pick("k_extract_tracks")}}
json.dump(traffic, open(os.path.join(P, "ncu_traffic.json"), "w"), indent=1)
d = json.loads(open(os.path.join(P, rnd + "_bench.json")).read().strip().splitlines()[-1])
print("value %.1f M/s, step %.3f ms, e2e %.1f M/s %.3f ms" % (d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e6,
                                                             d["e2e"]["ms_per_step"]))
print("stage_ms", {k: round(v, 3) for k, v in d["stage_ms"].items()})
for k in d["roofline_kernels"]:
    print("  %-46s %.3f ms  %5.0f GB/s  frac %.3f  traffic %s" % (k["kernel"], k["ms"], k["achieved"], k["frac"], k["traffic"]))
print("roofline k_ssp %.3f, stage %.3f, build %.3f" % (d["roofline"]["frac"], d["roofline_solve_stage"]["frac"],
                                                        d["roofline_edge_build"]["frac"]))
print("cpu one core %.0f det/s, text %.0f det/s" % (d["cpu_baseline"]["value"], d["cpu_baseline"]["text_interface"]["value"]))
r = json.loads(open(os.path.join(P, rnd + "_bench_reference_arm.json")).read().strip().splitlines()[-1])
print("reference arm %.0f det/s on %d cores" % (r["value"], r["cpu_baseline"]["cores"]))
