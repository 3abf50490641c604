#!/usr/bin/env python
"""bench.py -- tracked detections/s of the mot3d hot path (graph build + min-cost-flow solve + tracks) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--seqs S]

Workload (BASELINE.json configs[4], SURVEY.md 8d "C5 batched"): synthetic 2-D detections, 100 per frame, 1 000
frames per sequence, max_jump 8, cost = weight_distance_detections_2d(sigma_jump=1, sigma_distance=10,
max_distance=30, no histogram / bbox), weight_confidence(mul=1, bias=0), weight_source_sink=1, every sequence cut
into 20 non-overlapping windows of 50 frames (5 000 detections each = one graph).  A step = one pass of the hot path
over S sequences per GPU (S*20 graphs, S*100 000 detections), weak scaling over GPUs (different seeds per rank).

One JSON line on rank 0 (see the task's bench contract): value = device-resident throughput, e2e = through the
C-ABI host-buffer entry point (mot3d_track_batch_host: H2D + kernels + D2H inside the timed call), roofline for the
dominant kernel (+ roofline_edge_build), cpu_baseline timed on this box's host cores (N=1 only).
`--impl reference` times the reference CPU path instead: numpy port of build_graph's loops (oracle/) + the vendored
muSSP solver compiled from the reference's sources (oracle/_ref) when present, else the C port; all host cores.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FRAMES, PER_FRAME, WINDOW, MAX_JUMP = 1000, 100, 50, 8
DIST_KW = dict(sigma_jump=1, sigma_distance=10, max_distance=30, use_color_histogram=False, use_bbox=False)
METRIC = "tracked detections/sec (graph build + min-cost-flow solve + tracks)"


def gen_sequence(seed, frames=FRAMES, per_frame=PER_FRAME, extent=1000.0):
    """SURVEY.md 8(d) gen(F, P, seed)."""
    rng = np.random.RandomState(seed)
    pos = rng.rand(per_frame, 2) * extent
    vel = rng.randn(per_frame, 2) * 1.0
    out = np.empty((frames, per_frame, 2))
    for f in range(frames):
        pos = pos + vel + rng.randn(per_frame, 2) * 0.2
        out[f] = pos
    frame = np.repeat(np.arange(frames), per_frame)
    return frame, out.reshape(-1, 2)


def make_workload(n_seqs, seed0):
    frames, poss = [], []
    for s in range(n_seqs):
        f, p = gen_sequence(seed0 + s)
        frames.append(f)
        poss.append(p)
    frame = np.concatenate(frames)
    pos = np.concatenate(poss)
    per_win = WINDOW * PER_FRAME
    n = frame.shape[0]
    graph_ptr = np.arange(0, n + 1, per_win, dtype=np.int32)
    conf = np.full(n, 0.9)
    return graph_ptr, frame, pos, conf


# ------------------------------------------------------------------------------------------- clocks sampler
class ClockSampler(object):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def mark(self):
        """Samples from now on belong to the timed region (the process is started before the warm-up: nvidia-smi
        needs a few hundred ms before its first line)."""
        self.t0 = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        t1 = time.time()
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t0 = getattr(self, "t0", 0.0)
        inside = [r for t, r in self.rows if t0 <= t <= t1 + 0.03]
        if not inside and self.rows:  # region shorter than the sampling period: the sample closest to it
            inside = [min(self.rows, key=lambda tr: abs(tr[0] - 0.5 * (t0 + t1)))[1]]
        for r in inside:
            c = [x.strip() for x in r.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if c[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------- CPU reference arm
def _cpu_window(args):
    """One window on one core: numpy restatement of the build + muSSP solve (+ tracks). Returns detections done."""
    frame, pos, use_ref = args
    from oracle import oracle
    spec = oracle.CostSpec(weight_source_sink=1, max_jump=MAX_JUMP, mul=1, bias=0, **DIST_KW)
    n = len(frame)
    row_ptr, col, w = oracle.build_transitions(frame, pos, spec)
    tail, head, ww = oracle.graph_arcs(n, 1.0, oracle.node_weights(np.full(n, 0.9), spec), 0.0, row_ptr, col, w)
    if use_ref == "text":  # what mot3d itself does: graph_to_text lines -> wrapper.cpp's parse -> muSSP
        lines = oracle.arcs_to_text(2 * n + 2, tail, head, ww)
        res = oracle.mussp_solve_text(lines, len(tail))
    elif use_ref:
        res = oracle.mussp_solve(2 * n + 2, tail, head, np.round(ww, 7))
    else:
        res = oracle.ssp_solve(2 * n + 2, tail, head, oracle.quantise(ww))
    oracle.tracks_from_flow(n, tail, head, res["flow"])
    return n


def cpu_windows(n_windows, seed0):
    graph_ptr, frame, pos, _ = make_workload((n_windows + 19) // 20, seed0)
    out = []
    for g in range(n_windows):
        a, b = graph_ptr[g], graph_ptr[g + 1]
        out.append((frame[a:b].copy(), pos[a:b].copy()))
    return out


def run_reference(args):
    """--impl reference: the reference's CPU path on all host cores, bounded sample per step."""
    from oracle import oracle
    import multiprocessing as mp
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    oracle.build()
    use_ref = oracle.have_ref()
    cores = os.cpu_count() or 1
    per_step = max(cores, 8)  # one window per core and step
    wins = cpu_windows(per_step, 1000)
    tasks = [(f, p, use_ref) for f, p in wins]
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        for _ in range(max(1, min(args.warmup, 1))):
            pool.map(_cpu_window, tasks[:cores])
        t0 = time.time()
        done = 0
        for _ in range(args.steps):
            done += sum(pool.map(_cpu_window, tasks))
        dt = time.time() - t0
    value = done / dt
    kind = "reference" if use_ref else "port"
    sample = ("%d windows of 5000 detections per step, one per task over a %d-process pool: numpy port of build_graph "
              "(oracle/oracle.py) + %s + flow->tracks" %
              (per_step, cores, "vendored muSSP compiled from the reference zip (oracle/_ref), arrays in, no text"
               if use_ref else "C port of the solve (oracle/ssp_oracle.c)"))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "detections/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 weights / int64 costs",
            "data": "synthetic", "config": workload_config(per_step / 20.0, 1),
            "cpu_baseline": {"value": value, "unit": "detections/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "detections/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def workload_config(seqs, n_gpus):
    return {"workload": "C5 batched windows: synthetic 100 dets/frame, 1000 frames/sequence, max_jump=8, "
                        "20 windows x 50 frames (5000 detections) per sequence",
            "sequences_per_gpu": seqs, "graphs_per_gpu": int(seqs * 20), "detections_per_gpu": int(seqs * 100000),
            "max_jump": MAX_JUMP, "max_distance": 30, "parallelism": "graphs sharded over %d GPU(s), no data-path "
            "collective; track outputs gathered to rank 0 inside every step (one packed payload per rank, copy-engine put)" % n_gpus, "l2": "256 MiB L2 flush between timed iterations"}


# ------------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import mot3d_b200 as mot3d
    from mot3d_b200.pipeline import TrackPipeline

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; there is no CPU fallback (use --impl reference for the "
                           "CPU baseline)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    S = args.seqs
    graph_ptr, frame, pos, conf = make_workload(S, seed0=10000 * rank)
    n, G = len(frame), len(graph_ptr) - 1
    spec = mot3d.CostSpec(kind="detections_2d", distance=DIST_KW, confidence=dict(mul=1, bias=0),
                          weight_source_sink=1, max_jump=MAX_JUMP)
    spec_c = spec.to_c()
    host_batch, order = mot3d.DetectionBatch.from_arrays(graph_ptr, frame, pos, conf)
    dbatch = host_batch.to(dev)
    # size the arc buffers once (untimed): count pass through the public build
    gb = mot3d.build_graphs(dbatch, spec)
    E = gb.n_edges
    cap = int(E * 1.02) + 1024
    del gb
    pipe = TrackPipeline(n, G, cap, device=dev, fused=not args.unfused)
    STAGES, LAUNCHES_PER_RUN = pipe.stages, pipe.launches_per_run
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- the only communication of the path: the track outputs of every rank go to rank 0, inside every step, right
    # after the step's tracks on the launching stream (mot3d_b200.TrackGather: one fixed-size payload per rank, put
    # into rank 0's receive slot by the copy engines through a CUDA IPC peer buffer; MOT3D_BENCH_GATHER=collective
    # forces the torch.distributed gather of the same payload instead). Rank 0's step ends when every payload has
    # landed (flag wait on its stream).
    gather = None
    if world > 1:
        gather = mot3d.TrackGather(n, G, device=dev, transport=os.environ.get("MOT3D_BENCH_GATHER", "auto"))

    def gather_results():
        if gather is not None:
            gather.gather(dbatch.graph_ptr, pipe.track_count, pipe.track_ptr, pipe.track_det)
            gather.release()  # (the bench has no consumer on rank 0; the payloads are checked after the timed region)

    # ---- warm-up
    stream = torch.cuda.current_stream(dev)
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        pipe.run(dbatch, spec, spec_c)
        gather_results()
    barrier()
    pipe.check_overflow()

    # ---- timed: K steps, per-step CUDA events on the launching stream, L2 flushed between steps. The gather of a step
    #      runs inside the step's interval.
    step_ms = []
    stage_ms = np.zeros(len(STAGES))
    gather_ms = 0.0
    barrier()
    sampler.mark()
    wall0 = time.time()
    for k in range(args.steps):
        flush.fill_(k & 0xff)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(STAGES) + 1)]
        e_end = torch.cuda.Event(enable_timing=True)
        e_start = torch.cuda.Event(enable_timing=True)
        e_start.record(stream)
        pipe.run(dbatch, spec, spec_c, events=ev)
        gather_results()
        e_end.record(stream)
        e_end.synchronize()
        step_ms.append(e_start.elapsed_time(e_end))
        stage_ms += np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(len(STAGES))])
        gather_ms += ev[len(STAGES)].elapsed_time(e_end)
    barrier()
    wall = time.time() - wall0
    clocks = sampler.stop()
    total_ms = float(np.sum(step_ms))
    if world > 1:
        t = torch.tensor([total_ms, gather_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, gather_ms = float(t[0].item()), float(t[1].item())
    ms_per_step = total_ms / args.steps
    value = world * n / (ms_per_step * 1e-3)
    stage_ms /= args.steps
    gather_ms /= args.steps
    if gather is not None:
        # what arrived at rank 0 is what every rank sent (untimed): rank 0 compares each slot with the payload the
        # sender packs once more locally
        assert not gather.timed_out(), "track gather timed out"
        mine = torch.empty(2 * G + n, dtype=torch.int32, device=dev)
        mot3d._lib.check(mot3d._lib.lib().mot3d_pack_tracks(
            G, dbatch.graph_ptr.data_ptr(), pipe.track_count.data_ptr(), pipe.track_ptr.data_ptr(),
            pipe.track_det.data_ptr(), mine.data_ptr(), stream.cuda_stream))
        sent = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
        dist.gather(mine, sent, dst=0)
        if rank == 0:
            for r in range(world):
                assert torch.equal(gather.payload(r), sent[r]), "rank %d: payload at rank 0 differs from what was sent" % r
        gather_info = {"transport": gather.transport, "payload_bytes_per_rank": int(4 * (2 * G + n)),
                       "ms_per_step_max_over_ranks": gather_ms}
    else:
        gather_info = None

    # ---- kernel boundaries inside the solve stage (untimed extra steps; mot3d_solve_trace_events records CUDA events
    #      between the kernels of the calling thread's next solves): feeds the per-kernel roofline entries
    SOLVE_PARTS = ["k_first_tree", "k_components_tree (+offsets, fill)", "k_single_track + k_job_order",
                   "k_ssp_solve (one launch per band, overlapped)", "k_first_path_rule"]
    solve_part_ms = np.zeros(len(SOLVE_PARTS))
    tev = [torch.cuda.Event(enable_timing=True) for _ in range(len(SOLVE_PARTS) + 1)]
    for e in tev:
        e.record(stream)  # creates the underlying cudaEvent_t
    tev_arr = (ctypes.c_void_p * len(tev))(*[e.cuda_event for e in tev])
    TRACE_STEPS = 3
    for k in range(TRACE_STEPS):
        flush.fill_(k & 0xff)
        mot3d._lib.check(mot3d._lib.lib().mot3d_solve_trace_events(tev_arr, len(tev)))
        pipe.run(dbatch, spec, spec_c)
        mot3d._lib.check(mot3d._lib.lib().mot3d_solve_trace_events(None, 0))
        torch.cuda.synchronize(dev)
        solve_part_ms += np.array([tev[i].elapsed_time(tev[i + 1]) for i in range(len(SOLVE_PARTS))])
    solve_part_ms /= TRACE_STEPS

    # ---- end to end through the C-ABI host-buffer call (pinned host in, pinned host out)
    ht = mot3d.HostTracker(n_max=n, g_max=G, edge_capacity=cap, dim=2, device=dev)
    ht.load(host_batch)
    for _ in range(2):
        ht.run(spec_c)
    barrier()
    e2e_t = []
    for k in range(args.steps):
        flush.fill_(k & 0xff)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        ht.run(spec_c)  # returns after the D2H copies have landed
        e2e_t.append(time.perf_counter() - t0)
    e2e_s = float(np.sum(e2e_t))
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * n * args.steps / e2e_s

    # results of the two paths agree
    tracks_dev = pipe.tracks(graph_ptr)
    assert ht.tracks() == tracks_dev, "host-buffer path and device path disagree"
    assert (ht.h_total[:G].numpy() == pipe.total_cost.cpu().numpy()).all()

    if gather is not None:
        gather.close()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline accounting
    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_file):
        peak, peak_src = float(json.load(open(peaks_file))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    stats = pipe.stats.cpu().numpy().reshape(-1, mot3d._lib.SOLVE_STATS)[:G]
    # relaxation work actually done (k_ssp_solve only touches the branch it has to re-label):
    # every arc pull reads 4 B source + 8 B cost, every node pull moves 8 B dist + 8 B dist + 4 B parent
    solve_bytes = float(np.sum(stats[:, 3] * 12 + stats[:, 2] * 20))
    i_solve = STAGES.index("solve")
    solve_gbs = solve_bytes / (stage_ms[i_solve] * 1e-3) / 1e9
    # edge build (count + fill): detections 32 B (frame 4 + pos 16 + conf 8, padded) + row_ptr 4 B + arcs 12 B
    build_bytes = n * 32.0 + 4.0 * (n + 1) + E * 12.0
    build_stages = ["frame_sort", "edge_build"] if pipe.fused else ["frame_sort", "edge_count", "row_scan", "edge_fill"]
    build_ms = sum(stage_ms[STAGES.index(k)] for k in build_stages)
    build_gbs = build_bytes / (build_ms * 1e-3) / 1e9
    # measured DRAM traffic per launch: from the committed ncu capture of this very configuration, else null
    traffic_solve = traffic_build = None
    tf = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tf):
        t = json.load(open(tf))
        if t.get("seqs") == S:
            traffic_solve, traffic_build = t.get("k_ssp_solve"), t.get("edge_build")
    roofline_stage = {"bound": "hbm", "kernel": "solve stage",
                "kernels_in_launch_ms": "solve stage = k_first_tree + k_components_tree + k_single_track + k_job_order + "
                                        "k_ssp_solve (giants + main, overlapped) + k_first_path_rule",
                "achieved": solve_gbs, "peak": peak, "unit": "GB/s",
                "frac": solve_gbs / peak, "traffic": traffic_solve, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": solve_bytes, "launch_ms": float(stage_ms[i_solve]),
                "ascending_sweeps_per_graph_mean": float(stats[:, 0].mean()),
                "descending_sweeps_per_graph_mean": float(stats[:, 1].mean()),
                "branches_relabelled_per_graph_mean": float(stats[:, 11].mean()),
                "records_staged_per_graph_mean": float(stats[:, 12].mean()),
                "sweep_warp_cycles_per_step": {"ascending": float(stats[:, 13].sum() / max(stats[:, 15].sum(), 1)),
                                               "descending": float(stats[:, 14].sum() / max(stats[:, 16].sum(), 1))},
                "sweep_steps_per_graph_mean": {"ascending": float(stats[:, 15].mean()),
                                               "descending": float(stats[:, 16].mean())},
                "solver_block_busy_cycles": float(stats[0, 17]),
                "global_staging": {"components_per_graph_mean": float(stats[:, 22].mean()),
                                   "cycles_per_graph_mean": float(stats[:, 23].mean()),
                                   "ascending_cycles_per_step": float(stats[:, 18].sum() / max(stats[:, 20].sum(), 1)),
                                   "ascending_steps_per_graph_mean": float(stats[:, 20].mean()),
                                   "descending_cycles_per_hop": float(stats[:, 19].sum() / max(stats[:, 21].sum(), 1)),
                                   "descending_hops_per_graph_mean": float(stats[:, 21].mean())},
                "node_pulls_per_graph_mean": float(stats[:, 2].mean()),
                "arc_pulls_per_graph_mean": float(stats[:, 3].mean()),
                "phase_cycles_per_graph_mean": {k: float(stats[:, 4 + i].mean()) for i, k in enumerate(
                    ["sink_argmin", "list_branch", "stage_and_initial_labels", "walk", "fix_point", "certificate"])},
                "components_finished_by_certificate_per_graph_mean": float(stats[:, 10].mean()),
                "components_finished_by_single_track_pass_per_graph_mean": float(stats[:, 24].mean())}
    # the dominant kernel alone: k_ssp_solve (all band launches; they overlap), bytes = what its relabelling rounds and
    # its staging pulled (the first tree's share of the counters taken out), time = CUDA events around its launches
    ssp_bytes = max(solve_bytes - (12.0 * E + 40.0 * n), 0.0)
    ssp_ms = float(solve_part_ms[3])
    ssp_gbs = ssp_bytes / (ssp_ms * 1e-3) / 1e9 if ssp_ms > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": "k_ssp_solve", "achieved": ssp_gbs, "peak": peak, "unit": "GB/s",
                "frac": ssp_gbs / peak, "traffic": traffic_solve, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": ssp_bytes, "launch_ms": ssp_ms,
                "note": "latency bound by construction (dependent relabelling rounds in shared memory); "
                        "roofline_solve_stage has the whole stage, roofline_kernels every kernel"}
    roofline_build = {"bound": "hbm", "kernel": ("k_frame_sort + k_edge_build_fused" if pipe.fused else
                                 "k_frame_sort + k_edge_build_pruned<count> + scan + <fill> + k_edge_cost"),
                      "achieved": build_gbs, "peak": peak,
                      "unit": "GB/s", "frac": build_gbs / peak, "traffic": traffic_build,
                      "algorithmic_bytes_per_launch": build_bytes, "launch_ms": float(build_ms)}

    # ---- one entry per kernel (or back-to-back kernel group): live CUDA-event time, algorithmic bytes, DRAM traffic of
    #      the committed ncu capture of this configuration when there is one
    tk = {}
    if os.path.exists(tf):
        t = json.load(open(tf))
        if t.get("seqs") == S:
            tk = t.get("per_kernel", {})
    ft_bytes = 12.0 * E + 40.0 * n  # first tree: every arc (4 B source + 8 B cost) and every node (2 x 20 B) once
    per_kernel = []

    def kernel_entry(name, ms, nbytes, what):
        gbs = nbytes / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        per_kernel.append({"kernel": name, "ms": float(ms), "algorithmic_bytes": float(nbytes), "bytes_are": what,
                           "achieved": gbs, "unit": "GB/s", "frac": gbs / peak, "traffic": tk.get(name.split(" ")[0])})

    if pipe.fused:
        kernel_entry("k_frame_sort", stage_ms[STAGES.index("frame_sort")], n * (4 + 16 + 16 + 4.0),
                     "keys + positions in, x-sorted positions + permutation out")
        kernel_entry("k_edge_build_fused", stage_ms[STAGES.index("edge_build")], n * 28.0 + 4.0 * (n + 1) + E * 12.0,
                     "sorted detections in, row_ptr + 12 B per arc out")
    kernel_entry(SOLVE_PARTS[0], solve_part_ms[0], ft_bytes + 24.0 * n,
                 "12 B per arc + node costs in, distances / parent / branch labels out")
    kernel_entry(SOLVE_PARTS[1], solve_part_ms[1], 4.0 * E + 4.0 * (n + 1) + 16.0 * n + 64.0 * n,
                 "in-CSR structure + first tree in, permutation / inverse / component table + regrouped tree out")
    kernel_entry(SOLVE_PARTS[2], solve_part_ms[2], 36.0 * n, "first-tree labels in, tracks of single-track components out")
    kernel_entry(SOLVE_PARTS[3], solve_part_ms[3], max(solve_bytes - ft_bytes, 0.0),
                 "12 B per arc pull + 20 B per node pull (staging + relabelling rounds, counted by the kernel)")
    kernel_entry(SOLVE_PARTS[4], solve_part_ms[4], 8.0 * G, "path counts")
    kernel_entry("k_extract_tracks", stage_ms[STAGES.index("extract")], 20.0 * n, "next / prev in, tracks out")

    # ---- CPU baseline on this box (rank 0, N=1 only), bounded sample
    cpu_baseline = None
    if world == 1 and not args.no_cpu:
        from oracle import oracle
        oracle.build()
        use_ref = oracle.have_ref()
        wins = cpu_windows(args.cpu_windows, 1000)
        t0 = time.time()
        done = 0
        for f, p in wins:
            done += _cpu_window((f, p, use_ref))
        dt = time.time() - t0
        cpu_baseline = {"value": done / dt, "unit": "detections/s", "cores": 1,
                        "kind": "reference" if use_ref else "port",
                        "sample": "%d windows of 5000 detections, single thread: numpy port of build_graph "
                                  "(oracle/oracle.py) + %s + flow->tracks; host has %d cores" %
                                  (len(wins), "vendored muSSP from the reference zip (oracle/_ref)" if use_ref else
                                   "C port (oracle/ssp_oracle.c)", os.cpu_count() or 1),
                        "ms_per_window": 1e3 * dt / len(wins)}
        if use_ref:  # the same windows through the reference's text interface (mot3d/mot3d.py:218-226,
            # wrapper.cpp:39-88), fewer of them
            t0 = time.time()
            done_t = sum(_cpu_window((f, p, "text")) for f, p in wins[:8])
            dt_t = time.time() - t0
            cpu_baseline["text_interface"] = {"value": done_t / dt_t, "unit": "detections/s", "cores": 1,
                                              "sample": "%d of those windows: graph_to_text lines + wrapper.cpp's parse "
                                                        "+ muSSP" % min(8, len(wins)),
                                              "ms_per_window": 1e3 * dt_t / min(8, len(wins))}

    # ---- one graph at a time (SURVEY.md 8d metric ii): p50 / p99 per window on the reference's configurations, beside
    #      the reference's CPU path (tools/window_latency.py); N=1 only, bounded
    latency = None
    if world == 1 and not args.no_latency:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import window_latency
        latency = window_latency.measure(reps=args.latency_reps, budget_s=3.0, cpu=not args.no_cpu)

    line = {"metric": METRIC, "value": value, "unit": "detections/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64 weights / int64 costs", "data": "synthetic",
            "config": workload_config(S, world),
            "windows_per_s": world * G / (ms_per_step * 1e-3), "ms_per_window_amortised": ms_per_step / G,
            "stage_ms": {s: float(v) for s, v in zip(STAGES, stage_ms)},
            "e2e": {"value": e2e_value, "unit": "detections/s", "h2d_bytes_per_step": ht.h2d_bytes(),
                    "d2h_bytes_per_step": ht.d2h_bytes(), "ms_per_step": 1e3 * e2e_s / args.steps,
                    "api": "mot3d_track_batch_host (C ABI, pinned host buffers)"},
            "track_gather": gather_info,
            "gpu_launches": (LAUNCHES_PER_RUN + (2 if world > 1 else 0)) * args.steps, "roofline": roofline, "roofline_solve_stage": roofline_stage, "roofline_edge_build": roofline_build,
            "roofline_kernels": per_kernel,
            "cpu_baseline": cpu_baseline, "latency": latency, "clocks": clocks, "wall_s_timed_region": wall,
            "n_arcs_per_gpu": int(E), "paths_per_graph_mean": float(pipe.n_paths.cpu().numpy()[:G].mean())}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--seqs", type=int, default=60, help="sequences (of 20 windows) per GPU and step")
    ap.add_argument("--cpu-windows", type=int, default=40, help="windows in the single-thread CPU baseline sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-latency", action="store_true", help="skip the per-window latency section")
    ap.add_argument("--latency-reps", type=int, default=40)
    ap.add_argument("--unfused", action="store_true", help="count -> scan -> fill graph build instead of the "
                    "single-pass kernel (device-resident arm only)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
