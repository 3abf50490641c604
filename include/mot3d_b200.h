/* mot3d_b200 -- C ABI of the B200-native mot3d tracking hot path (graph build + min-cost-flow solve + tracks).
 *
 * This is the drop-in boundary: it replaces the reference's only native interface,
 *     wrapper.solve(graph_as_text: list[str], verbose: int) -> list[(tail, head)]
 *     (/root/reference/mot3d/solvers/wrappers/muSSP/wrapper.cpp:171, 319-324; called from mot3d/mot3d.py:160-161),
 * and moves the Python work either side of it (build_graph's loops, mot3d/mot3d.py:75-141; graph_to_text, :218-226;
 * path enumeration, :165-184) behind the same boundary.  Plain pointers and sizes only; no torch types.
 *
 * Conventions
 *   - every function returns MOT3D_OK (0) or a negative MOT3D_E_* code; mot3d_last_error() gives a thread-local
 *     message.  Nothing is printed, nothing throws across the ABI.
 *   - unless a name ends in _host, every pointer is a DEVICE pointer owned by the caller (e.g. tensor.data_ptr());
 *     the library allocates no user-visible memory.  `stream` is a cudaStream_t passed as void*; calls are
 *     asynchronous with respect to the host.
 *   - a batch is a concatenation of independent graphs (windows / views / sequences): graph g owns detections
 *     [graph_ptr[g], graph_ptr[g+1]).  Inside a graph detections are in the reference's node order: stable sort
 *     by frame index (mot3d/mot3d.py:61-64); detection m of a graph is pre-node 2+2m / post-node 3+2m there.
 *   - costs are int64 in units of 1e-7, the precision the reference hands to muSSP ("{:0.7f}", mot3d/mot3d.py:224).
 *   - no global mutable state: safe from several host threads on different streams/devices.
 */
#ifndef MOT3D_B200_H
#define MOT3D_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MOT3D_OK 0
#define MOT3D_E_INVALID (-1)     /* bad argument */
#define MOT3D_E_CUDA (-2)        /* a CUDA call failed (message has the CUDA error string) */
#define MOT3D_E_CAPACITY (-3)    /* caller-provided buffer too small */
#define MOT3D_E_UNSUPPORTED (-4) /* valid in the reference, not built here (message says what) */

#define MOT3D_MAX_JUMP 256
#define MOT3D_INF_COST ((int64_t)1 << 60) /* "no arc" for entry/exit costs */

#define MOT3D_DIR_IN 0  /* CSR rows = arc heads (pre_j), columns = tails (post_i), ascending: what the solver pulls from */
#define MOT3D_DIR_OUT 1 /* CSR rows = arc tails (post_i), columns = heads (pre_j), ascending: graph_to_text order */

#define MOT3D_SOLVE_STATS 25 /* int64 per graph written to mot3d_solve_batch's stats_out */

#define MOT3D_FLAG_ENTRY 1 /* Detection.entry_points (mot3d/types.py:24, mot3d/mot3d.py:90-91) */
#define MOT3D_FLAG_EXIT 2  /* Detection.exit_point   (mot3d/types.py:25, mot3d/mot3d.py:98-99) */

/* Packed detections, structure of arrays (replaces the Detection2D/Detection3D objects of mot3d/types.py:11-45). */
typedef struct mot3d_dets {
    int64_t n;            /* detections in the whole batch */
    int32_t dim;          /* 2 (Detection2D) or 3 (Detection3D) */
    int32_t hist_channels; /* C, 0 when hist == NULL */
    int32_t hist_bins;    /* B */
    int32_t reserved;
    const int32_t *tkey;  /* [n] time keys made by mot3d_make_keys: frame index shifted per graph so that the batch is
                             one non-decreasing sequence and keys of different graphs differ by more than max_jump */
    const double *pos;    /* [dim][n] planar positions */
    const double *conf;   /* [n] Detection.confidence */
    const uint8_t *flags; /* [n] MOT3D_FLAG_*; NULL = entry and exit allowed everywhere (the reference default) */
    const double *bbox;   /* [4][n] planar xmin,ymin,xmax,ymax; NULL unless use_bbox */
    const double *hist;   /* [n][C][B] colour histograms (mean-subtracted, mot3d/distance_functions.py:10-19); NULL unless use_hist */
    double *sorted_pos;   /* optional scratch [dim][n]: positions sorted by x inside every frame, made by mot3d_frame_sort */
    int32_t *sorted_perm; /* optional scratch [n]: original index of every sorted entry. When both are set,
                             mot3d_edge_count / mot3d_edge_fill search x-windows instead of testing every pair
                             (same arcs, same order); NULL = brute force */
} mot3d_dets;

/* Parameters of the built-in costs (mot3d/weight_functions.py:13-17,37-38,43-47) and of build_graph
 * (mot3d/mot3d.py:20-23), pre-digested by the host so that the device reproduces the reference's float64 values:
 *   sq_dist_max      largest s with sqrt(s) <= max_distance in IEEE double (the reference tests dist > max_distance)
 *   sigma_*_sq       sigma**2 as the reference evaluates it
 *   neg_exp_jump[k]  -exp(-(k)*sigma_jump) for k = jump-1 = 0..max_jump-1, evaluated by the host's numpy exp */
typedef struct mot3d_cost_spec {
    int32_t max_jump;
    int32_t use_bbox;
    int32_t use_hist;
    int32_t defer_cost; /* 1: the build stores the distance factor alone in weight_out and no costs (as with use_hist);
                           an appearance pass (mot3d_edge_appearance_views) finishes the arcs */
    double sq_dist_max;
    double sigma_distance_sq;
    double sigma_hist_sq;
    double sigma_bbox_sq;
    double conf_mul;
    double conf_bias;
    int64_t entry_cost; /* quantised weight_source_sink */
    int64_t exit_cost;  /* 0 in the reference (mot3d/mot3d.py:99) */
    double neg_exp_jump[MOT3D_MAX_JUMP];
} mot3d_cost_spec;

const char *mot3d_version(void);
const char *mot3d_last_error(void);

/* Process-wide test hooks / tuning knobs (no environment variables are read). value = NaN puts the built-in default
 * back. Names: "host_chunks", "host_pieces", "host_split", "host_edge_weight", "host_trace", "host_no_pack"
 * (mot3d_track_batch_host: how the batch is cut; timings to stderr; unpacked track_ptr), "edge_cap", "edge_acap"
 * (mot3d_edge_build: staging capacities, 0 forces the global-memory paths), "solve_no_fast" (no single-track pass),
 * "big_delta" (bucket width of solve_big.cuh in cost units). MOT3D_E_INVALID for an unknown name. */
int32_t mot3d_set_option(const char *name, double value);

/* ---- graph build: replaces build_graph's loops (mot3d/mot3d.py:75-141) + graph_to_text (:218-226) ------------ */

/* tkey[i] = frame[i] - frame[first of graph] + base(graph), base = running sum of (frame span + max_jump + 1).
 * graph_base_ws: scratch of n_graphs int64. Preconditions every later stage relies on are checked on the way and
 * reported by OR-ing bits into *error_flags (device int32, optional; the same word mot3d_edge_build's overflow_flag
 * uses, so one read-back covers the build): MOT3D_ERR_UNSORTED = the frames of a graph are not non-decreasing,
 * MOT3D_ERR_KEY_RANGE = the keys of the batch do not fit 31 bits (sum of frame spans too large: split the batch). */
#define MOT3D_ERR_ARC_CAPACITY 1 /* mot3d_edge_build / mot3d_edge_fill: more arcs than edge_capacity */
#define MOT3D_ERR_WINDOW 2       /* mot3d_edge_build: more than 2^24 detections inside one time window */
#define MOT3D_ERR_UNSORTED 4
#define MOT3D_ERR_KEY_RANGE 8
int32_t mot3d_make_keys(const int32_t *frame, const int32_t *graph_ptr, int32_t n_graphs, int32_t max_jump,
                        int32_t *tkey_out, int64_t *graph_base_ws, int32_t *error_flags, void *stream);

/* per-detection arc costs: entry (S->pre), node (pre->post, weight_confidence_detections, weight_functions.py:37-38),
 * exit (post->T).  MOT3D_INF_COST where flags forbid the arc. */
int32_t mot3d_node_costs(const mot3d_dets *dets, const mot3d_cost_spec *spec, int64_t *entry_cost_out,
                         int64_t *node_cost_out, int64_t *exit_cost_out, void *stream);

/* fills dets->sorted_pos / dets->sorted_perm (both must be set): per frame, detections ranked by (x, index). */
int32_t mot3d_frame_sort(const mot3d_dets *dets, void *stream);

/* pass 1 of the windowed all-pairs kernel: row_count[r] = number of transition arcs of row r
 * (pairs with 0 < frame[j]-frame[i] <= max_jump and dist <= max_distance; mot3d/mot3d.py:122-127). */
int32_t mot3d_edge_count(const mot3d_dets *dets, const mot3d_cost_spec *spec, int32_t direction, int32_t *row_count_out,
                         void *stream);

size_t mot3d_scan_workspace_bytes(int64_t n);
/* out[0] = 0, out[i+1] = in[0] + ... + in[i]  (n+1 outputs). */
int32_t mot3d_exclusive_scan_i32(const int32_t *in, int64_t n, int32_t *out, void *ws, size_t ws_bytes, void *stream);

/* pass 2: columns (ascending), float64 weights (NULL to skip) and quantised costs of every arc.
 * When spec->use_hist is set, `weight_out` is required and receives the distance factor only; finish with
 * mot3d_edge_appearance.  Arcs beyond edge_capacity are not written and *overflow_flag (device int32) is set to 1. */
int32_t mot3d_edge_fill(const mot3d_dets *dets, const mot3d_cost_spec *spec, int32_t direction, const int32_t *row_ptr,
                        int64_t edge_capacity, int32_t *col_out, int64_t *cost_out, double *weight_out,
                        int32_t *overflow_flag, void *stream);

/* Single-pass build (csrc/edge_fused.cu): rows [row_begin, row_end) of the transition CSR in ONE launch - what
 * mot3d_edge_count + mot3d_exclusive_scan_i32 + mot3d_edge_fill produce, bit for bit, without the second sweep over
 * the candidates, the scan kernels and the (column, squared distance) round trip through HBM. Needs the x-sorted
 * copies (mot3d_frame_sort). Arc offsets come from a decoupled look-back over 256-row tiles:
 *   row_ptr_out[r] for r in [row_begin, row_end] (absolute offsets into col_out / cost_out / weight_out)
 *   arc_base_in   device int64 or NULL: offset of the first arc of row_begin (NULL = 0), so that consecutive calls over
 *                 consecutive row ranges append to one CSR;  arc_total_out: device int64 or NULL, receives the offset
 *                 after the last arc (may alias arc_base_in)
 *   overflow_flag device int32: bit 0 set when arcs beyond edge_capacity were dropped (row_ptr_out stays exact), bit 1
 *                 when a tile's time window holds more than 2^24 detections (not supported)
 *   ws            mot3d_edge_build_workspace_bytes(row_end - row_begin) bytes of device scratch
 * weight_out / use_hist as in mot3d_edge_fill. */
size_t mot3d_edge_build_workspace_bytes(int64_t n_rows);
int32_t mot3d_edge_build(const mot3d_dets *dets, const mot3d_cost_spec *spec, int32_t direction, int64_t row_begin,
                         int64_t row_end, int32_t *row_ptr_out, int64_t edge_capacity, int32_t *col_out,
                         int64_t *cost_out, double *weight_out, int32_t *overflow_flag, const int64_t *arc_base_in,
                         int64_t *arc_total_out, void *ws, size_t ws_bytes, void *stream);

/* colour-histogram factor (color_histogram_similarity, mot3d/distance_functions.py:21-34) multiplied in, then the
 * bbox and jump factors, then quantisation. One warp per row. arc_capacity = entries of col / weight / cost: rows that
 * reach beyond it (a build that ran out of capacity and said so) are clipped. */
int32_t mot3d_edge_appearance(const mot3d_dets *dets, const mot3d_cost_spec *spec, int32_t direction,
                              const int32_t *row_ptr, const int32_t *col, int64_t arc_capacity, double *weight_inout,
                              int64_t *cost_out, void *stream);

/* The 2-D views of 3-D detections (Detection3D.detections_2d, mot3d/types.py:40-45): detection i owns the entries
 * [view_ptr[i], view_ptr[i+1]) in the iteration order of its dict. */
typedef struct mot3d_views {
    int64_t n;                /* detections */
    int64_t n_entries;        /* view_ptr[n] */
    int32_t hist_channels;
    int32_t hist_bins;
    const int32_t *view_ptr;  /* [n+1] */
    const int32_t *view_id;   /* [n_entries] camera label */
    const double *bbox;       /* [4][n_entries] planar xmin,ymin,xmax,ymax of the 2-D detection; NULL unless use_bbox */
    const double *hist;       /* [n_entries][C][B] its colour histogram; NULL unless use_hist */
} mot3d_views;

/* Appearance terms of weight_distance_detections_3d (mot3d/weight_functions.py:55-71): (1 - NCC) and (1 - bbox
 * similarity) averaged over the views both detections were seen in (iterated in the tail detection's dict order, as
 * the reference does), multiplied into the distance factor left by a build with spec->defer_cost, then the jump
 * factor and the quantisation. One warp per row.
 *   spec        the real spec (use_hist / use_bbox as the user gave them)
 *   error_flag  device int32, set to 1 when two linked detections share no view: the reference's np.mean([]) is nan
 *               there and the text it feeds muSSP undefined; the arc gets weight nan and cost 0 */
int32_t mot3d_edge_appearance_views(const mot3d_dets *dets, const mot3d_views *views, const mot3d_cost_spec *spec,
                                    int32_t direction, const int32_t *row_ptr, const int32_t *col,
                                    int64_t arc_capacity, double *weight_inout, int64_t *cost_out,
                                    int32_t *error_flag, void *stream);

/* ---- tracklet linking (second pass): replaces build_graph on DetectionTracklet lists (mot3d/mot3d.py:106-141)
 *      with weight_distance_tracklets_2d/_3d (mot3d/weight_functions.py:80-160) and linear_motion
 *      (mot3d/distance_functions.py:47-63) ------------------------------------------------------------------- */

/* One graph of n tracklets in the caller's order (the reference does not sort tracklets). Every tracklet brings the
 * window of detections at its head (most recent end) and at its tail (oldest end): DetectionTracklet2D/3D,
 * mot3d/types.py:94-184. */
typedef struct mot3d_tracklets {
    int32_t n;
    int32_t dim;
    int32_t hist_channels;
    int32_t hist_bins;
    int64_t n_head_points;        /* head_ptr[n] */
    int64_t n_tail_points;        /* tail_ptr[n] */
    const int32_t *head_ptr;      /* [n+1] */
    const int32_t *tail_ptr;      /* [n+1] */
    const int32_t *head_idx;      /* frame index of every head-window point, ascending inside a tracklet */
    const int32_t *tail_idx;
    const double *head_pos;       /* [dim][n_head_points] planar */
    const double *tail_pos;       /* [dim][n_tail_points] planar */
    const uint8_t *tail_access_point; /* [n] tail.access_point (2-D only) or NULL */
    const uint8_t *head_has_hist; /* [n] head.color_histogram is not None; NULL = all have one */
    const uint8_t *tail_has_hist;
    const double *head_hist;      /* [n][C][B] median histogram of the head window; NULL unless use_hist */
    const double *tail_hist;
    double *head_fit_ws;          /* scratch [n][dim][2]: line fits, written by mot3d_tracklet_edge_count */
    double *tail_fit_ws;
    /* 3-D tracklets with use_hist (weight_distance_tracklets_3d, mot3d/weight_functions.py:146-158): the per-view 2-D
     * tracklets of DetectionTracklet3D.tracklets_2d (mot3d/types.py:142-184), entries of tracklet i =
     * [view_ptr[i], view_ptr[i+1]) in the dict's order. NULL = single-histogram mode (head_hist / tail_hist). */
    const int32_t *view_ptr;      /* [n+1] */
    const int32_t *view_id;       /* [view_ptr[n]] camera label */
    const double *view_head_hist; /* [view_ptr[n]][C][B] median histogram at the head of the per-view tracklet */
    const double *view_tail_hist; /* ... at its tail */
} mot3d_tracklets;

typedef struct mot3d_tracklet_spec {
    int32_t max_jump;
    int32_t use_hist;
    int32_t check_access_point; /* 1 for weight_distance_tracklets_2d (:101-102), 0 for _3d */
    int32_t reserved;
    double sq_dist_max;         /* largest squared head-tail distance accepted; negative = max_distance is None */
    double sigma_motion_sq;
    double sigma_hist_sq;
    double alpha;
    double cutoff_motion;
    double cutoff_appearance;
} mot3d_tracklet_spec;

/* rows = tail tracklet t1, columns = head tracklet t2 (ascending), i.e. the reference's arc order.
 * overflow_flag: bit 0 = arcs beyond edge_capacity dropped; bit 1 = (view mode) two linked tracklets share no view: the
 * reference's np.mean([]) is nan there, the arc gets weight nan and cost 0 */
int32_t mot3d_tracklet_edge_count(const mot3d_tracklets *trk, const mot3d_tracklet_spec *spec, int32_t *row_count_out,
                                  void *stream);
int32_t mot3d_tracklet_edge_fill(const mot3d_tracklets *trk, const mot3d_tracklet_spec *spec, const int32_t *row_ptr,
                                 int64_t edge_capacity, int32_t *col_out, int64_t *cost_out, double *weight_out,
                                 int32_t *overflow_flag, void *stream);

/* ---- connected components: every graph is split into groups of detections no arc connects; the solver treats each
 *      group as an independent small graph (same optimum: S and T are the only coupling). No reference counterpart.
 *   comp_perm_out[q]   detection at regrouped position q (graph g keeps the positions [graph_ptr[g], graph_ptr[g+1]))
 *   comp_inv_out[d]    regrouped position of detection d
 *   cgraph_ptr_out     [n_total + n_graphs] component c = positions [cgraph_ptr[c], cgraph_ptr[c+1])
 *   cgraph_graph_out   [n_total] graph of component c;  graph_cbase_out [n_graphs] first component of graph g
 *   n_cgraphs_out      device int32[4]: [0] = number of components, [1] = 0 (work-queue cursor of the solver) */
size_t mot3d_components_workspace_bytes(int64_t n_total, int32_t n_graphs);
int32_t mot3d_components(int32_t n_graphs, int64_t n_total, const int32_t *graph_ptr, const int32_t *in_ptr,
                         const int32_t *in_src, int64_t arc_capacity, int32_t *comp_perm_out, int32_t *comp_inv_out,
                         int32_t *cgraph_ptr_out, int32_t *cgraph_graph_out, int32_t *graph_cbase_out,
                         int32_t *n_cgraphs_out, void *ws, size_t ws_bytes, void *stream);

/* ---- solve: replaces wrapper.solve (wrapper.cpp:171-317) and the muSSP core it drives ------------------------ */

size_t mot3d_solve_workspace_bytes(int64_t n_total, int32_t n_graphs, int64_t arc_capacity);

/* Successive shortest paths on the residual network of every graph: each graph is split into its connected components,
 * single-track components are finished from the first tree, the others are solved one per thread block from a work queue.
 *   in_ptr/in_src/in_cost : MOT3D_DIR_IN CSR of the transition arcs over the whole batch (in_src = global indices);
 *                   arc_capacity = entries present in in_src/in_cost: a graph whose rows reach beyond it (arc buffer
 *                   overflow during the build) is skipped and gets n_paths = -1
 *   next_out[i]  : successor of detection i on its track (global index), -2 if the track ends at i, -1 if i is unused
 *   prev_out[i]  : predecessor, -2 if a track starts at i, -1 if unused
 *   n_paths_out[g], total_cost_out[g] : K and the sum of accepted path costs
 *   path_cost_out : optional [n_total]; the accepted path costs of graph g are the entries < MOT3D_INF_COST of its slice
 *                   [graph_ptr[g], graph_ptr[g+1]) (grouped by connected component, ascending inside a component)
 *   stats_out     : optional int64 [n_graphs][MOT3D_SOLVE_STATS]: [0..3] = ascending sweeps, descending sweeps, node
 *                   pulls, arc pulls (what bench.py's roofline accounting needs); [4..9] = SM cycles spent in:
 *                   sink arg-min, branch listing, staging + initial labels, path walk, fix-point rounds, optimality
 *                   certificate; [10] = components finished by the certificate;
 *                   [11] = branches re-labelled, [12] = records listed for them, [13..14] = SM cycles of the sweep
 *                   warp in ascending / descending sweeps, [15..16] = ascending / descending steps, [17] (first graph of the
 *                   batch only) = cycles all solver blocks were busy, summed; [18..21] = [13..16] of the
 *                   components solved on the global staging area alone, [22] = how many, [23] = their cycles; [24] =
 *                   components finished by the single-track pass (k_single_track) without entering the solver
 * Acceptance rule of wrapper.cpp:199-267 on integers: path 1 always, then while cost < 0. */
int32_t mot3d_solve_batch(int32_t n_graphs, int64_t n_total, const int32_t *graph_ptr, int32_t max_graph_nodes,
                          const int32_t *tkey, int64_t arc_capacity,
                          const int32_t *in_ptr, const int32_t *in_src, const int64_t *in_cost,
                          const int64_t *entry_cost, const int64_t *node_cost, const int64_t *exit_cost,
                          int32_t *next_out, int32_t *prev_out, int32_t *n_paths_out, int64_t *total_cost_out,
                          int64_t *path_cost_out, int64_t *stats_out, void *ws, size_t ws_bytes, void *stream);

/* mot3d_solve_batch in phases, for a caller whose graphs become ready in pieces (mot3d_track_batch_host does this with
 * the pieces of its H2D pipeline): MOT3D_SOLVE_BEGIN once, MOT3D_SOLVE_FRONT
 * for every range of graphs [g_first, g_first + g_count) whose in-CSR rows are complete (connected components + first
 * shortest-path tree: per-graph kernels, a third of the stage), MOT3D_SOLVE_BACK once when every graph has been through
 * FRONT (component numbering, single-track pass, work queue, successive shortest paths, first-path rule). Phases may be
 * OR-ed; all calls take the same arrays; calls on different streams must be ordered by the caller (events). */
#define MOT3D_SOLVE_BEGIN 1
#define MOT3D_SOLVE_FRONT 2
#define MOT3D_SOLVE_BACK 4
int32_t mot3d_solve_batch_phase(int32_t phases, int32_t g_first, int32_t g_count, int32_t n_graphs, int64_t n_total,
                                const int32_t *graph_ptr, int32_t max_graph_nodes, const int32_t *tkey,
                                int64_t arc_capacity, const int32_t *in_ptr, const int32_t *in_src, const int64_t *in_cost,
                                const int64_t *entry_cost, const int64_t *node_cost, const int64_t *exit_cost,
                                int32_t *next_out, int32_t *prev_out, int32_t *n_paths_out, int64_t *total_cost_out,
                                int64_t *path_cost_out, int64_t *stats_out, void *ws, size_t ws_bytes, void *stream);

/* path_cost_out of mot3d_solve_batch -> per graph the accepted path costs ascending (the order muSSP finds them,
 * wrapper.cpp:199-267) at the front of the graph's slice, MOT3D_INF_COST behind them. In place; scratch: int64 [n_total].
 * Quadratic in the paths of one graph: a graph with more than max_paths of them is left as it is and *too_many_flag
 * (device int32, caller clears it) is set. */
int32_t mot3d_order_path_costs(int32_t n_graphs, const int32_t *graph_ptr, int64_t *path_cost, int64_t *scratch,
                               int32_t max_paths, int32_t *too_many_flag, void *stream);

/* Measurement hook: the calling thread's next mot3d_solve_batch calls record events[k] (cudaEvent_t, may be NULL) on
 * their stream at the kernel boundaries of the solve: 0 = start, 1 = after the connected components, 2 = after the first
 * tree, 3 = after the single-track pass and the queue order, 4 = after the successive-shortest-path launches, 5 = end.
 * events = NULL switches it off. The array must stay alive while the hook is set. */
#define MOT3D_SOLVE_TRACE_EVENTS 6
int32_t mot3d_solve_trace_events(void **events, int32_t n_events);

/* ---- trajectories: replaces _run_ssp's tail (mot3d/mot3d.py:165-184) ---------------------------------------- */

/* Tracks of graph g, ordered by their first detection (= the reference's DFS order over SOURCE's successors):
 *   track_count_out[g] = K
 *   track_ptr_out[graph_ptr[g] + g + k] (k = 0..K) = start of track k inside the graph's slice of track_det_out
 *   track_det_out[graph_ptr[g] + ...]   = detection indices LOCAL to the graph (0-based position in sorted order) */
int32_t mot3d_extract_tracks(int32_t n_graphs, const int32_t *graph_ptr, const int32_t *next, const int32_t *prev,
                             int32_t *track_count_out, int32_t *track_ptr_out, int32_t *track_det_out, void *stream);

/* ---- track gather (multi-GPU): the only communication of the path ----------------------------------------------
 * Whole graphs are sharded over the GPUs of a box; every rank sends its tracks to rank 0 as ONE
 * fixed-size int32 payload, so that no size has to reach the host before the copy is issued:
 *     payload = [ K_g (G words) | L_g (G words) | track_det (n words) ],   mot3d_track_payload_words(n, G) words
 * K_g = tracks of graph g, L_g = detections on them; the first detection of every track carries
 * MOT3D_TRACK_START_BIT, which makes the offsets of mot3d_extract_tracks redundant (mot3d_unpack_tracks rebuilds
 * them). It stands in for the Python lists solve_graph returns (mot3d/mot3d.py:208-211) between processes. */
#define MOT3D_TRACK_START_BIT 0x80000000u
#define MOT3D_PEER_HANDLE_BYTES 64
int64_t mot3d_track_payload_words(int64_t n_total, int32_t n_graphs);
int32_t mot3d_pack_tracks(int32_t n_graphs, const int32_t *graph_ptr, const int32_t *track_count,
                          const int32_t *track_ptr, const int32_t *track_det, int32_t *payload_out, void *stream);
int32_t mot3d_unpack_tracks(int32_t n_graphs, const int32_t *graph_ptr, const int32_t *payload,
                            int32_t *track_count_out, int32_t *track_ptr_out, int32_t *track_det_out, void *stream);
/* Peer buffers: device memory allocated by the library (cudaMalloc, zeroed) that other processes of the box can
 * open through CUDA IPC (the 64-byte handle travels by any host channel, e.g. torch.distributed's object broadcast).
 * create/destroy in the owning process, open/close in the others. These four calls synchronise with the device. */
int32_t mot3d_peer_buffer_create(size_t bytes, void **dev_ptr_out, unsigned char handle_out[MOT3D_PEER_HANDLE_BYTES]);
int32_t mot3d_peer_buffer_open(const unsigned char handle[MOT3D_PEER_HANDLE_BYTES], void **dev_ptr_out);
int32_t mot3d_peer_buffer_close(void *dev_ptr);
int32_t mot3d_peer_buffer_destroy(void *dev_ptr);
/* put: `bytes` from src (local) to peer_dst (opened peer buffer) by the copy engines, then `step` into *peer_flag
 * (stream order: the flag lands after the payload). local_flag_scratch = one int32 of local device memory.
 * wait (owner): a one-thread kernel on `stream` that returns once flags[r * stride_words] - step >= 0 for every
 * r in [1, world); after ~2 s it gives up and writes the offending rank to *timeout_flag (device, 0 = fine). */
int32_t mot3d_peer_put(void *peer_dst, const void *src, size_t bytes, int32_t *peer_flag, int32_t *local_flag_scratch,
                       int32_t step, void *stream);
int32_t mot3d_peer_wait(const int32_t *flags, int32_t stride_words, int32_t world, int32_t step, int32_t *timeout_flag,
                        void *stream);
/* flow control of a double-buffered slot: the owner publishes "consumed up to step s" with set_word on its own buffer,
 * a sender waits with wait_word (one-thread kernel reading the peer word over NVLink until *word - need >= 0; gives
 * up after ~2 s and sets *timeout_flag) before it overwrites the slot of step s + 2. */
int32_t mot3d_peer_set_word(int32_t *word, int32_t value, void *stream);
int32_t mot3d_peer_wait_word(const int32_t *word, int32_t need, int32_t *timeout_flag, void *stream);


/* ---- input-side prep: merge_close_points (mot3d/utils/filt.py:6-24; examples/soldiers_3d/example.py:29-31) ---- */

/* Per point set s = points [set_ptr[s], set_ptr[s+1]) (one frame): clusters = connected components of "distance <=
 * eps" (DBSCAN with min_samples=1), numbered by their first member; every cluster replaced by the mean of its members.
 *   points            [N][dim] row-major (as users hold them), dim <= 3;  max_set_points = largest set (<= 8192)
 *   label_out[i]      cluster of point i inside its set
 *   cluster_count_out [n_sets]
 *   merged_out        [N][dim]: the means of set s are rows set_ptr[s] .. set_ptr[s] + cluster_count_out[s] - 1 */
int32_t mot3d_merge_close_points(const int32_t *set_ptr, int32_t n_sets, int32_t max_set_points, int32_t dim,
                                 const double *points, double eps, int32_t *label_out, int32_t *cluster_count_out,
                                 double *merged_out, void *stream);

/* ---- device-resident glue between the detection pass and the tracklet pass (one graph) ------------------------ */

/* Tracks of pass 1 (mot3d_extract_tracks on ONE graph of n detections) -> the tracklets of pass 2, without leaving the
 * device: split_trajectory_modulo(length) + remove_short_trajectories(th_length) (mot3d/utils/trajectory.py:216-248) and
 * the head / tail windows of DetectionTracklet2D/3D(tracklet, window) (mot3d/types.py:94-184; window 0 = None).
 * Tracklets come out ordered by the frame they start in (key_out, the solver's time key of pass 2: an arc only leads to
 * a tracklet that starts after its tail ends, so the order is topological); the reference keeps them in track order -
 * the linked trajectories are the same, their numbering is not.
 *   trk_first_out[k], trk_len_out[k] : tracklet k = entries [first, first + len) of track_det
 *   head_ptr_out / tail_ptr_out [n+1], head_idx_out / tail_idx_out [n], head_pos_out / tail_pos_out [dim][n] planar with
 *   stride n: the arrays of struct mot3d_tracklets, with n_head_points = n_tail_points = n there
 *   n_tracklets_out : device int32 */
size_t mot3d_tracklets_from_tracks_workspace_bytes(int64_t n);
int32_t mot3d_tracklets_from_tracks(int32_t n, int32_t dim, const int32_t *frame, const double *pos,
                                    const int32_t *track_count, const int32_t *track_ptr, const int32_t *track_det,
                                    int32_t split_length, int32_t th_length, int32_t window, int32_t *n_tracklets_out,
                                    int32_t *trk_first_out, int32_t *trk_len_out, int32_t *key_out, int32_t *head_ptr_out,
                                    int32_t *tail_ptr_out, int32_t *head_idx_out, int32_t *tail_idx_out,
                                    double *head_pos_out, double *tail_pos_out, void *ws, size_t ws_bytes, void *stream);

/* The same for the online loop (projects/PETS2009S2L1/singleview_tracking.py:94-121): the tracklet set is the n_carry
 * trajectories carried over from earlier batches (Scene.active, whole, in id order) followed by the new pieces, as in the
 * reference's list; ranks are by (start frame, list position). carry_ptr [n_carry + 1] / carry_frame [points] /
 * carry_pos [dim][carry_stride] planar describe the carried trajectories (device arrays, typically the output of the
 * previous mot3d_online_concat). head_pos_out / tail_pos_out are planar with stride out_stride (>= the points any
 * window set can hold: carried points + n). item_out [n_carry + n]: per rank, c < n_carry = carried trajectory c, else
 * n_carry + number of the new piece in track order; for a carried trajectory trk_first_out addresses carry_frame.
 * span_out (may be NULL) [2][..] interleaved: first and last frame of the tracklet at every rank.
 * n = 0 (a batch without detections) is allowed: track_count must then point at a zero. */
int32_t mot3d_online_tracklets(int32_t n, int32_t dim, const int32_t *frame, const double *pos,
                               const int32_t *track_count, const int32_t *track_ptr, const int32_t *track_det,
                               int32_t split_length, int32_t th_length, int32_t window, int32_t n_carry,
                               const int32_t *carry_ptr, const int32_t *carry_frame, const double *carry_pos,
                               int64_t carry_stride, int64_t out_stride, int32_t *n_tracklets_out, int32_t *item_out,
                               int32_t *trk_first_out, int32_t *trk_len_out, int32_t *key_out, int32_t *head_ptr_out,
                               int32_t *tail_ptr_out, int32_t *head_idx_out, int32_t *tail_idx_out, double *head_pos_out,
                               double *tail_pos_out, int32_t *span_out, void *ws, size_t ws_bytes, void *stream);

/* The online loop's write-back: concat_tracklets + interpolate_trajectory(position, bbox)
 * (mot3d/utils/trajectory.py:11-69) for n_out output trajectories at once. Trajectory t is the chain of tracklets
 * chain[chain_ptr[t] .. chain_ptr[t+1]) (ranks of mot3d_online_tracklets, in time order) and occupies points
 * [out_ptr[t], out_ptr[t+1]) of the new arrays: one point per frame from its first to its last, skipped frames filled
 * by v1 + (v2 - v1) / (t2 - t1) * (t - t1). bbox / carry_bbox / new_bbox are planar [4][stride] like the positions.
 * hist_len > 0: the points' colour histograms travel too (hist [n][hist_len] of this batch or NULL, carry_hist /
 * carry_has of the carried points, new_hist [points][hist_len] / new_has out); filled points have none.
 * *error_flag |= 1 if a chain does not cover its range exactly (wrong out_ptr, tracklets out of order). */
int32_t mot3d_online_concat(int32_t n_out, const int32_t *chain_ptr, const int32_t *chain, const int32_t *out_ptr,
                            const int32_t *item, const int32_t *trk_first, const int32_t *trk_len, int32_t n_carry,
                            int32_t n, int32_t dim, const int32_t *track_det, const int32_t *frame, const double *pos,
                            const double *bbox, const int32_t *carry_frame, const double *carry_pos,
                            const double *carry_bbox, int64_t carry_stride, int32_t *new_frame, double *new_pos,
                            double *new_bbox, int64_t new_stride, int32_t hist_len, const double *hist,
                            const double *carry_hist, const uint8_t *carry_has, double *new_hist, uint8_t *new_has,
                            int32_t *error_flag, void *stream);

/* Median colour histogram of the head window and of the tail window of every tracklet (ranks of
 * mot3d_online_tracklets / mot3d_tracklets_from_tracks, same `window`): DetectionTracklet2D's head.color_histogram /
 * tail.color_histogram, np.median over the detections of the window that have one (mot3d/types.py:69-86,109-125).
 * hist [n][hist_len] are the histograms of this batch's detections (NULL = none has one); carry_hist / carry_has those
 * of the carried points (points filled by interpolation have none). Outputs [n_tracklets][hist_len] and the has-flags
 * struct mot3d_tracklets takes. item may be NULL when n_carry = 0. */
int32_t mot3d_tracklet_window_hist(int32_t n_tracklets, const int32_t *item, const int32_t *trk_first,
                                   const int32_t *trk_len, int32_t n_carry, int32_t window, int32_t hist_len,
                                   const int32_t *track_det, const int32_t *frame, const double *hist,
                                   const int32_t *carry_frame, const double *carry_hist, const uint8_t *carry_has,
                                   double *head_hist_out, double *tail_hist_out, uint8_t *head_has_out,
                                   uint8_t *tail_has_out, void *stream);

/* Out-CSR of mot3d_tracklet_edge_fill (rows = arc tails) -> the solver's inputs for one graph of *n_nodes_dev nodes
 * (<= n_nodes_max): in-CSR, constant entry / node / exit costs, graph_ptr = {0, n}. ws: int32 [n_nodes_max + 1]. */
int32_t mot3d_tracklet_graph_in_csr(const int32_t *n_nodes_dev, int32_t n_nodes_max, const int32_t *row_ptr,
                                    const int32_t *col, const int64_t *cost, int64_t entry_cost, int64_t node_cost,
                                    int64_t exit_cost, int32_t *in_ptr_out, int32_t *in_src_out, int64_t *in_cost_out,
                                    int64_t *entry_out, int64_t *node_out, int64_t *exit_out, int32_t *graph_ptr_out,
                                    int32_t *ws, void *stream);

/* Chains of tracklets (mot3d_extract_tracks of pass 2) -> chains of detections (concat_tracklets,
 * mot3d/utils/trajectory.py:11-12): out_ptr [K + 1], out_det = indices of pass-1 detections. */
int32_t mot3d_concat_tracklets(const int32_t *track_count, const int32_t *track_ptr, const int32_t *track_trk,
                               const int32_t *trk_first, const int32_t *trk_len, const int32_t *track_det_pass1,
                               int32_t *out_count, int32_t *out_ptr, int32_t *out_det, void *stream);

/* ---- input-side prep: color_histogram (mot3d/distance_functions.py:10-19) ------------------------------------- */

/* Histograms of image patches, all boxes of an image in one launch, written straight into the `hist` array of
 * mot3d_dets. For box b the patch is image[ymin:ymax, xmin:xmax] (numpy slice semantics: bounds clipped to the image):
 * 8-bit RGB -> CIE Lab exactly as cv2.cvtColor(patch, cv2.COLOR_RGB2LAB) on uint8 (integer tables, bit-exact), then per
 * requested channel np.histogram(values, bins=n_bins, range=(0, 255)) minus its mean.
 *   image_rgb    device uint8 [height][row_stride_bytes], pixels interleaved R, G, B
 *   boxes        device int32 [n_boxes][4] = xmin, ymin, xmax, ymax
 *   channels     HOST int32 [n_channels], Lab channel per output row (reference default (0, 1, 2))
 *   bin_of_value device uint16 [256]: histogram bin of every 8-bit value; may be NULL for n_bins == 255 (the reference's
 *                default: bin v for v < 255, value 255 joins bin 254)
 *   hist_out     device float64 [n_boxes][n_channels][n_bins] */
int32_t mot3d_color_histograms(const uint8_t *image_rgb, int32_t height, int32_t width, int64_t row_stride_bytes,
                               const int32_t *boxes, int32_t n_boxes, const int32_t *channels, int32_t n_channels,
                               int32_t n_bins, const uint16_t *bin_of_value, double *hist_out, void *stream);

/* ---- one call with HOST buffers: pack -> H2D -> build -> solve -> tracks -> D2H ------------------------------ */

typedef struct mot3d_host_batch {
    int32_t n_graphs;
    int32_t dim;
    int64_t n;                 /* total detections */
    const int32_t *graph_ptr;  /* host [n_graphs+1] */
    const int32_t *frame;      /* host [n], sorted inside each graph */
    const double *pos;         /* host [dim][n] */
    const double *conf;        /* host [n] */
    const uint8_t *flags;      /* host [n] or NULL */
    const double *bbox;        /* host [4][n] or NULL */
} mot3d_host_batch;

typedef struct mot3d_host_result {
    int32_t *track_count; /* host [n_graphs] */
    int32_t *track_ptr;   /* host [n + n_graphs] */
    int32_t *track_det;   /* host [n] */
    int64_t *total_cost;  /* host [n_graphs] */
    int64_t n_edges;      /* out: transition arcs built */
    int64_t h2d_bytes;    /* out: bytes the call copied host -> device */
    int64_t d2h_bytes;    /* out: bytes it copied device -> host (with one chunk only the K + 1 meaningful entries of
                             every graph's track_ptr slice travel; the rest of track_ptr is left untouched) */
} mot3d_host_result;

size_t mot3d_track_batch_workspace_bytes(int64_t n, int32_t n_graphs, int64_t edge_capacity);
/* Synchronous on `stream` (returns after the results are in host memory). Host buffers should be pinned. */
int32_t mot3d_track_batch_host(const mot3d_host_batch *batch, const mot3d_cost_spec *spec, mot3d_host_result *result,
                               void *device_ws, size_t ws_bytes, int64_t edge_capacity, void *stream);

/* mot3d_track_batch_host keeps a few streams, events and small pinned buffers per calling host thread and device,
 * created on first use. A thread that is done with the call may release them (every device it used); the next call
 * creates them again. Nothing of the caller's may be in flight. */
int32_t mot3d_host_release(void);

#ifdef __cplusplus
}
#endif
#endif /* MOT3D_B200_H */
