"""Print the numbers of a bench.py JSON line that matter while tuning: python tools/bench_summary.py FILE"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value %.1f M det/s  step %.3f ms  e2e %.1f M (%.3f ms)" % (d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e6,
                                                                 d["e2e"]["ms_per_step"]))
print("stage_ms", {k: round(v, 3) for k, v in d["stage_ms"].items()})
r = d["roofline"]
for k in ["ascending_sweeps_per_graph_mean", "branches_relabelled_per_graph_mean", "records_staged_per_graph_mean",
          "sweep_warp_cycles_per_step", "node_pulls_per_graph_mean", "arc_pulls_per_graph_mean",
          "phase_cycles_per_graph_mean", "solver_block_busy_cycles", "frac"]:
    if k in r:
        print(k, r[k])
g = r.get("global_staging", {})
print("dbg", g)
