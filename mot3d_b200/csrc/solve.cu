// Min-cost-flow solve: successive shortest paths on the residual network, one thread block per connected component.
//
// Replaces wrapper.solve (reference mot3d/solvers/wrappers/muSSP/wrapper.cpp:171-317) and the muSSP core it drives
// (zip!/muSSP-master/muSSP/Graph.cpp).  Same network, same path-update rule, same acceptance rule, exact int64:
//   network        S -> pre_i (entry), pre_i -> post_i (node cost), post_i -> T (exit), post_i -> pre_j (transition);
//                  unit capacities                                   (mot3d/mot3d.py:75-141)
//   first tree     one relaxation pass over the frame layers         (Graph::shortest_path_dag, Graph.cpp:139-187)
//   components     a graph falls apart into connected components (csrc/components.cu); each one is a job of the work
//                  queue, staged ONCE into the shared memory of a block (csrc/solve_sm.cuh), solved path after path
//                  without global traffic, and only its flow is written back. The queue is cut into bands by the
//                  bytes a component needs; every band is drained by a launch whose blocks have just that much
//                  shared memory, so that an SM holds as many components in flight as fit. Components that fit no
//                  band or whose costs need more than 32 bits are solved on a staging area in global memory
//                  (csrc/solve_global.cuh).
//   next paths     muSSP's observation (Graph.cpp:341-343, 394-503): after a path through S -> pre_a is flipped,
//                  only the branch of the shortest-path tree that hung from that arc loses its labels; every
//                  other node keeps its exact distance.  The branch is re-labelled by exact Bellman-Ford relaxations
//                  in pull mode (synchronous rounds in shared memory, alternating layer sweeps on the global staging
//                  area).  No tolerances, no heap.
//   path update    reverse the arcs of the path                      (Graph::flip_path, Graph.cpp:266-335); with unit
//                  capacities the residual state is just next/prev per detection
//   acceptance     path 1 always; path k >= 2 while its cost < 0     (wrapper.cpp:199-267 on integers)
// Flow read-out (make_output2, wrapper.cpp:139-169) is unnecessary: next/prev are the flow.
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <type_traits>

#include "common.cuh"

namespace m3 {

#ifndef M3_LPN
#define M3_LPN 4
#endif
constexpr int LPN = M3_LPN;  // lanes cooperating on one node's in-arcs (staging, relabelling rounds)

// state markers (slots / positions are >= 0)
constexpr int NONE = -1;  // detection unused
constexpr int TERM = -2;  // next: track ends here (post->T saturated); prev: track starts here (S->pre saturated)
constexpr int PAR_S = -1;     // pre reached from S
constexpr int PAR_POST = -2;  // pre reached from its own post (reversed node arc)
constexpr int PAR_NONE = -3;
constexpr int DL_PRE = 1 << 29;  // branch flags of a record: its pre / post node is being re-labelled
constexpr int DL_POST = 1 << 30;

// costs of a component staged in shared memory are kept in 32 bits (k_first_tree checks per graph that they fit:
// win_wide); NO_ARC32 = no such arc
constexpr int NO_ARC32 = 0x7fffffff;
__device__ __forceinline__ bool fits_narrow(int64_t v) {
    return v >= INF_CUT || (v == (int64_t)(int)v && (int)v != NO_ARC32);
}
struct GRec;
struct StArc;
constexpr int SM_BANDS = 3;  // bands of the work queue solved in shared memory (+ one on the global staging area)

struct SolveParams {
    int32_t n_graphs;
    int32_t g_first;  // k_first_tree: block b works on graph g_first + b
    int32_t rcap;    // shared-memory staging area of this launch: most records a component may have ...
    int32_t sbytes;  // ... and its size in bytes
    int32_t role;   // queue band this launch drains (SM_BANDS = the global staging area's)
    int32_t chain;  // bit 0: let the next launch start early, bit 1: wait for the launch before (see k_ssp_solve)
    // queue bands: a component goes to the first band whose staging area holds it (records and bytes, see
    // staging_bytes), else to the band of the global staging area
    int32_t band_rec[SM_BANDS];
    int32_t band_bytes[SM_BANDS];
    int32_t *comp_class;  // [n_total] per component: band * JOB_CLASSES + size class (k_single_track -> k_job_order)
    int64_t n_total;
    const int32_t *graph_ptr;
    const int32_t *tkey;
    const int32_t *in_ptr;
    const int32_t *in_src;
    const int64_t *in_cost;
    const int64_t *en;
    const int64_t *cn;
    const int64_t *ex;
    int32_t *next_out;
    int32_t *prev_out;
    int32_t *n_paths;
    int64_t *total;
    int64_t *path_cost;    // optional [n_total], indexed like the regrouped order (component's k-th path at c0 + k)
    int64_t *stats;        // optional [n_graphs][MOT3D_SOLVE_STATS], see include/mot3d_b200.h
    int64_t arc_capacity;  // arcs actually present in in_src/in_cost (graphs reaching beyond it are skipped)
    // component graphs (csrc/components.cu): detections regrouped by connected component
    const int32_t *comp_perm;   // [n_total] detection at every regrouped position
    const int32_t *comp_inv;    // [n_total] regrouped position of every detection
    const int32_t *cgraph_ptr;  // [n_cgraphs + 1] regrouped positions
    const int32_t *cgraph_win;  // [n_cgraphs] graph of every component
    const int32_t *win_cbase;   // [n_graphs] first component of every graph
    int32_t *win_wide;          // [n_graphs] 1 = some cost does not fit 32 bits (k_first_tree)
    int4 *job_tab;              // [n_cgraphs] the work queue, largest components first (k_job_order): x = first
                                // position, y = detections (0 = nothing to do), z = graph, w = 1 if the costs are wide
    int32_t *n_cgraphs;         // device: [0] = number of component graphs, [1] = work-queue cursor
    int32_t *job_meta;          // device: [0 .. 3*JOB_CLASSES) components per size class and band (k_single_track),
                                // then as many scatter cursors (k_job_order), then JOB_Q x: length of the queue, end
                                // of band 2, end of band 1, cursor of the launch for band 2 / 1 / 0
    int32_t fast;               // 1 = k_single_track may finish single-track components
    int64_t *first_cost;        // [n_total] per component: cost of its first path when that was not kept (>= 0)
    int32_t *first_sink;        // [n_total] ... and the position where it ends
    // (the first shortest-path tree lives in ft / ft_tmp, see FtTmp)
    int32_t *ws_lvl;       // [n_total + n_graphs] frame layers
    FtTmp *ft_tmp;         // [n_total] what k_first_tree leaves, in detection order (k_components_tree regroups it)
    FtTmp *ft;             // [n_total] ... and in the regrouped order, parents / branch labels as positions
    int32_t *ft_anc;       // [n_total] the branch labels of ft_tmp once more (k_components_tree starts from them)
    // global staging area (components that do not fit shared memory; spilled arc lists of those that do)
    GRec *ws_rec;          // [n_total]
    int32_t *ws_tk;        // [n_total] layer keys
    int32_t *ws_bl;        // [n_total] x 3 slot lists
    int32_t *ws_hop;
    int32_t *ws_pl;
    StArc *ws_arc;         // [arc_capacity + n_total + 8]: the list of detection d starts at in_ptr[d] + d
    // staging area of solve_big.cuh (components beyond the shared-memory bands, 32-bit costs)
    longlong2 *big_res_e;  // [n_total]
    int4 *big_res_f;       // [n_total]
    unsigned *big_listed;  // [n_total / 16 + n_total + 8]
    longlong2 *big_old_e;  // [n_total]
    long long *big_prio;   // [n_total]
    int *big_far;          // [2 * n_total]
    long long big_delta;   // width of a bucket of label increases
    int32_t *big_arc;      // 3 x [big_arc_stride]: cost, source, out-arc head; blocks handed out through big_arc_next
    int64_t big_arc_stride;
    unsigned long long *big_arc_next;
};

__device__ __forceinline__ KeyIdx group_min(KeyIdx best, unsigned mask) {
#pragma unroll
    for (int d = 1; d < LPN; d <<= 1) {
        KeyIdx o = shfl_xor_ki(best, d, mask);
        if (o.key < best.key || (o.key == best.key && o.idx >= 0 && (best.idx < 0 || o.idx < best.idx))) best = o;
    }
    return best;
}

// Block-wide exclusive scan with ONE barrier: every warp publishes its total, every thread sums the totals of the
// warps before it. `scratch` must hold 2 * 32 ints; `phase` alternates 0/1 between consecutive calls.
__device__ __forceinline__ int block_excl_scan1(int v, int *scratch, int phase, int *total) {
    const int incl = warp_incl_scan(v);
    const int w = warp_id(), l = lane_id();
    const int nw = (blockDim.x + 31) >> 5;
    int *sc = scratch + 32 * phase;
    if (l == 31) sc[w] = incl;
    __syncthreads();
    int base = 0, tot = 0;
    for (int k = 0; k < nw; k++) {
        const int t = sc[k];
        if (k < w) base += t;
        tot += t;
    }
    *total = tot;
    return base + incl - v;
}

// Block-wide sum; scratch must hold 33 ints. Two barriers.
__device__ __forceinline__ int block_sum(int v, int *scratch) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    __syncthreads();
    if (lane_id() == 0) scratch[warp_id()] = v;
    __syncthreads();
    int tot = 0;
    for (int k = 0; k < (int)((blockDim.x + 31) >> 5); k++) tot += scratch[k];
    __syncthreads();
    return tot;
}
constexpr int BIG_MIN_WIDTH = 48;  // detections per frame layer from which solve_big.cuh takes a component

extern __shared__ __align__(16) unsigned char smem_raw[];

// ---- first shortest-path tree of every graph (before it is split into components): no flow yet, the network is a
//      DAG, one pull sweep over the frame layers with the whole block (Graph::shortest_path_dag, Graph.cpp:139-187).
//      State is written in detection order (FtTmp; k_components_tree moves it to the regrouped order); parent and
//      branch label (the detection whose S arc starts the node's tree path) are global detection indices.
constexpr int FT_W = 2048;  // ring of post / branch labels (power of two)
template <int FT_LPN>
__global__ void __launch_bounds__(256, 5) k_first_tree(const SolveParams p) {
    __shared__ int s_scan[33];
    __shared__ int64_t s_dp[FT_W];
    __shared__ int s_an[FT_W];
    const int g = blockIdx.x + p.g_first;
    const int g0 = p.graph_ptr[g];
    const int n = p.graph_ptr[g + 1] - g0;
    const int T = blockDim.x, tid = threadIdx.x;
    if (tid == 0) {
        p.n_paths[g] = 0;
        p.total[g] = 0;
        if (p.stats)
            for (int k = 0; k < MOT3D_SOLVE_STATS; k++) p.stats[(size_t)MOT3D_SOLVE_STATS * g + k] = 0;
    }
    if (n <= 0) return;
    const int32_t *tkey = p.tkey + g0;
    const int32_t *in_ptr = p.in_ptr + g0;
    // first_cost / first_sink are indexed by component, and the components of a graph may be numbered anywhere in
    // [0, n_total): the whole batch's arrays are cleared by mot3d_solve_batch before this launch
    for (int i = tid; i < n; i += T)
        if (p.path_cost) p.path_cost[g0 + i] = INF;
    if ((int64_t)in_ptr[n] > p.arc_capacity) {  // arc buffer overflow (reported by the fill kernel): skipped
        if (tid == 0) p.n_paths[g] = -1;
        for (int i = tid; i < n; i += T) {
            p.next_out[g0 + i] = NONE;
            p.prev_out[g0 + i] = NONE;
        }
        return;
    }
    int32_t *lvl = p.ws_lvl + g0 + g;
    int L = 0;
    for (int base = 0; base < n; base += T) {
        const int i = base + tid;
        const int f = (i < n && (i == 0 || tkey[i] != tkey[i - 1])) ? 1 : 0;
        int tot;
        const int ex_ = block_excl_scan(f, s_scan, &tot);
        if (f) lvl[L + ex_] = i;
        L += tot;
    }
    if (tid == 0) lvl[L] = n;
    __syncthreads();
    // FT_LPN lanes per node, every lane keeps FT_U independent arc loads in flight: a layer of ~100 nodes with ~8
    // in-arcs each is one pass of the block. Arcs only reach back max_jump frames, so the post labels and branch
    // labels of the last FT_W detections live in a ring in shared memory: the chain a layer waits for is row offsets ->
    // arcs -> shared memory (two global levels instead of four); a source that has left the ring (very dense frames)
    // is read from global memory. Sources of one detection sit in the same component, where regrouped positions
    // ascend with the detection index: ties go to the lowest detection index = lowest position, and the parent is kept
    // as a detection index until the end of the sweep.
    constexpr int FT_U = 4;
    const int sub = tid % FT_LPN, grp = tid / FT_LPN, ngrp = T / FT_LPN;
    int wide = 0;  // some cost needs more than 32 bits: the packed staging of k_ssp_solve is off
    for (int l = 0; l < L; l++) {
        const int a = lvl[l], b = lvl[l + 1];
        const int ring_lo = g0 + b - FT_W;  // global index of the oldest detection this layer may read from the ring
        for (int jb = a; jb < b; jb += ngrp) {
            const int j = jb + grp;  // detection (local)
            const bool valid = j < b;
            KeyIdx best;
            best.key = INF;
            best.idx = PAR_NONE;
            int64_t e_ = INF, c_ = 0, x_ = INF;
            int64_t w_best = INF, w_min = INF;  // cost of the parent arc / of the cheapest in-arc (k_single_track)
            if (valid) {
                const int e0 = in_ptr[j], deg = in_ptr[j + 1] - e0;
                if (sub == 0) {  // the node's own costs: in flight together with the arc loads
                    e_ = p.en[g0 + j];
                    c_ = p.cn[g0 + j];
                    x_ = p.ex[g0 + j];
                }
                for (int kb = 0; kb < deg; kb += FT_LPN * FT_U) {
                    int src[FT_U];
                    int64_t w[FT_U], di[FT_U];
#pragma unroll
                    for (int u = 0; u < FT_U; u++) {
                        const int k = kb + u * FT_LPN + sub;
                        src[u] = k < deg ? p.in_src[e0 + k] : -1;
                        w[u] = k < deg ? p.in_cost[e0 + k] : 0;
                    }
#pragma unroll
                    for (int u = 0; u < FT_U; u++) {
                        if (src[u] < 0)
                            di[u] = INF;
                        else if (src[u] >= ring_lo)
                            di[u] = s_dp[(src[u] - g0) & (FT_W - 1)];
                        else
                            di[u] = p.ft_tmp[src[u]].dpost;
                    }
#pragma unroll
                    for (int u = 0; u < FT_U; u++) {
                        wide |= (w[u] != (int64_t)(int)w[u]) ? 1 : 0;
                        if (src[u] >= 0 && w[u] < w_min) w_min = w[u];
                        if (di[u] < INF_CUT) {
                            const int64_t cand = di[u] + w[u];
                            if (cand < best.key || (cand == best.key && src[u] < best.idx)) {
                                best.key = cand;
                                best.idx = src[u];
                                w_best = w[u];
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int d = 1; d < FT_LPN; d <<= 1) {  // lowest source wins ties
                const KeyIdx o = shfl_xor_ki(best, d);
                const int64_t ow = __shfl_xor_sync(0xffffffffu, w_best, d);
                const int64_t om = __shfl_xor_sync(0xffffffffu, w_min, d);
                if (o.key < best.key || (o.key == best.key && o.idx >= 0 && (best.idx < 0 || o.idx < best.idx))) {
                    best = o;
                    w_best = ow;
                }
                w_min = om < w_min ? om : w_min;
            }
            if (valid && sub == 0) {
                int anc = -1;
                if (best.idx >= 0)
                    anc = best.idx >= ring_lo ? s_an[(best.idx - g0) & (FT_W - 1)] : p.ft_tmp[best.idx].anc;
                wide |= (fits_narrow(e_) && fits_narrow(x_) && c_ == (int64_t)(int)c_) ? 0 : 1;
                if (e_ < best.key) {
                    best.key = e_;
                    best.idx = PAR_S;
                    anc = g0 + j;
                }
                // what solve_component's optimality certificate asks of this detection if its tree arc ends up
                // matched: no negative entry / exit cost; the arc is the cheapest in-arc, costs mc <= 0, cn + mc <= 0
                int cert = ((e_ < INF_CUT && e_ < 0) || (x_ < INF_CUT && x_ < 0)) ? 0 : 1;
                if (best.idx >= 0 && !(w_best <= 0 && c_ + w_best <= 0 && w_min >= w_best)) cert = 0;
                int64_t dpost = INF;
                if (best.key < INF_CUT) {
                    dpost = best.key + c_;
                } else {  // unreachable
                    best.key = INF;
                    best.idx = PAR_NONE;
                    anc = -1;
                }
                FtTmp t;
                t.dpre = best.key;
                t.dpost = dpost;
                t.par = best.idx;
                t.anc = anc;
                t.cert = cert;
                t.pad = 0;
                p.ft_tmp[g0 + j] = t;
                p.ft_anc[g0 + j] = anc;
                s_dp[j & (FT_W - 1)] = dpost;
                s_an[j & (FT_W - 1)] = anc;
            }
        }
        __syncthreads();
    }
    wide = __syncthreads_or(wide);
    if (tid == 0) p.win_wide[g] = wide;
    if (p.stats && tid == 0) {
        int64_t *o = p.stats + (size_t)MOT3D_SOLVE_STATS * g;
        o[0] = 1;
        o[2] = 2 * (int64_t)n;
        o[3] = in_ptr[n] - in_ptr[0];
    }
}

#include "solve_global.cuh"
#include "solve_sm.cuh"
#include "solve_big.cuh"

// ---- single-track components. Three out of four components of a crowded window are one object seen once per frame:
//      the first shortest path runs through all of its detections in order, the optimality certificate of
//      solve_component (matched arcs are the cheapest in-arcs of their heads, no negative entry / exit cost) holds,
//      and the component is finished with K = 1. That can be read off the first tree without staging anything: one
//      warp per component checks parent chain, sink arg-min and the certificate bits k_first_tree left per detection
//      (a few coalesced loads per lane, no arc is read again) and writes the track; k_job_order then leaves the component out of the solver's queue. The answer is exactly what
//      solve_component would return (same sink by its tie-breaking, same acceptance rule); anything that does not
//      qualify is left to it. first_sink[c] = FAST_DONE marks a finished component.
constexpr int JOB_CLASSES = 64;
__device__ __forceinline__ int job_class(int n) {
    if (n < 4) return n;
    const int lg = 31 - __clz(n);
    const int c = 4 * (lg - 1) + ((n >> (lg - 2)) & 3);  // 4 -> 4, 8 -> 8, ..., 8192 -> 48
    return c < JOB_CLASSES ? c : JOB_CLASSES - 1;
}
constexpr int FAST_DONE = -2;
// Work queue: SM_BANDS bands of components solved in shared memory (the staging area grows from band to band, so
// fewer of them are in flight per SM), then the band of what fits nowhere or has costs beyond 32 bits (global staging
// area). Queue order: that last band first, then the shared-memory bands from the largest down; inside a band by size
// class, largest first (a component is solved by one block, path after path: longest-processing-time-first keeps the
// tail short).
constexpr int JOB_BANDS = SM_BANDS + 1;
constexpr int JOB_NB = JOB_BANDS * JOB_CLASSES;
constexpr int JOB_Q = 2 * JOB_NB;                         // offset of the queue descriptor in job_meta:
constexpr int JOB_META_INTS = JOB_Q + 2 * JOB_BANDS + 2;  // [0 .. JOB_BANDS] band starts in queue order, then cursors
__device__ __forceinline__ int band_class(const SolveParams &p, int n, long long arcs, bool wide) {
    const long long need = staging_bytes(n, arcs);
    int band = SM_BANDS;
    if (!wide)
        for (int b = SM_BANDS - 1; b >= 0; b--)
            if (n <= p.band_rec[b] && need <= p.band_bytes[b]) band = b;
    return job_class(n) + band * JOB_CLASSES;
}
__device__ __forceinline__ void single_track(const SolveParams &p, const int c);
__global__ void __launch_bounds__(256) k_single_track(const SolveParams p) {
    // the number of components lives on the device: a fixed grid of warps strides over them
    const int n_comp = p.n_cgraphs[0];
    const int warps = (int)((gridDim.x * blockDim.x) >> 5);
    for (int c = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5); c < n_comp; c += warps) single_track(p, c);
}

__device__ __forceinline__ bool single_track_try(const SolveParams &p, const int c, const int c0, const int n,
                                                 const int g);
// finished here, or counted into its size class for the solver's queue (k_job_order)
__device__ __forceinline__ void single_track(const SolveParams &p, const int c) {
    const int c0 = p.cgraph_ptr[c], n = p.cgraph_ptr[c + 1] - c0, g = p.cgraph_win[c];
    if (n <= 0 || p.n_paths[g] < 0) return;  // empty, or the graph was skipped (arc buffer overflow)
    if (p.fast && n <= 4096 && single_track_try(p, c, c0, n, g)) return;
    // the queue band depends on the bytes the component needs in a staging area: count its in-arcs
    long long arcs = 0;
    for (int i = lane_id(); i < n; i += 32) {
        const int d = p.comp_perm[c0 + i];
        arcs += p.in_ptr[d + 1] - p.in_ptr[d];
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) arcs += __shfl_xor_sync(0xffffffffu, arcs, d);
    if (lane_id() == 0) {
        const int b = band_class(p, n, arcs, p.win_wide[g] != 0);
        p.comp_class[c] = b;
        atomicAdd(&p.job_meta[b], 1);
    }
}

__device__ __forceinline__ bool single_track_try(const SolveParams &p, const int c, const int c0, const int n,
                                                 const int g) {
    const int lane = lane_id();
    // the first tree is the chain c0 -> c0+1 -> ..., every detection passes the certificate (k_first_tree left one bit
    // per detection, computed from the arcs it was reading anyway)
    bool ok = true;
    int64_t key_last = INF, key_rest = INF;
    for (int i = lane; i < n && ok; i += 32) {
        const int q = c0 + i;
        const FtTmp t = p.ft[q];
        const int par = t.par;
        ok = ((i == 0) ? (par == PAR_S) : (par == q - 1)) && t.cert != 0;
        const int64_t ex = p.ex[p.comp_perm[q]];
        const int64_t dpost = t.dpost;
        const int64_t key = (ex < INF_CUT && dpost < INF_CUT) ? dpost + ex : INF;
        if (i == n - 1)
            key_last = key;
        else if (key < key_rest)
            key_rest = key;
    }
    if (!__all_sync(0xffffffffu, ok)) return false;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        key_last = min(key_last, __shfl_xor_sync(0xffffffffu, key_last, d));
        key_rest = min(key_rest, __shfl_xor_sync(0xffffffffu, key_rest, d));
    }
    // the sink arg-min of solve_component takes the lowest slot among equal keys: the last detection has to win
    // strictly, and the path has to pay (wrapper.cpp:263-267)
    if (!(key_last < key_rest) || key_last >= 0) return false;
    // (3) the track
    for (int i = lane; i < n; i += 32) {
        const int d = p.comp_perm[c0 + i];
        p.next_out[d] = (i + 1 < n) ? p.comp_perm[c0 + i + 1] : TERM;
        p.prev_out[d] = (i > 0) ? p.comp_perm[c0 + i - 1] : TERM;
    }
    if (lane == 0) {
        atomicAdd(&p.n_paths[g], 1);
        atomicAdd((unsigned long long *)&p.total[g], (unsigned long long)key_last);
        if (p.path_cost) p.path_cost[c0] = key_last;
        p.first_sink[c] = FAST_DONE;
        if (p.stats) atomicAdd((unsigned long long *)(p.stats + (size_t)MOT3D_SOLVE_STATS * g + 24), 1ull);
    }
    return true;
}

// ---- work queue order: largest components first (longest-processing-time-first keeps the tail of the solver short:
//      a component is solved by one block, path after path). One block: counting sort by size class, 4 classes per
//      power of two; the order inside a class does not matter.
__global__ void __launch_bounds__(256) k_job_order(const SolveParams p) {
    // The class counts were taken by k_single_track; every block turns them into class starts and scatters its share
    // of the components.
    __shared__ int s_start[JOB_NB];
    const int n_comp = p.n_cgraphs[0];
    const int32_t *hist = p.job_meta;
    int32_t *cursor = p.job_meta + JOB_NB;
    for (int c = threadIdx.x; c < JOB_NB; c += blockDim.x) {
        int run = 0;
        for (int b = JOB_NB - 1; b > c; b--) run += hist[b];
        s_start[c] = run;
        if (blockIdx.x == 0) {
            int32_t *q = p.job_meta + JOB_Q;
            if (c % JOB_CLASSES == JOB_CLASSES - 1) {  // everything above class c = the bands above c's
                const int k = JOB_BANDS - 1 - c / JOB_CLASSES;  // position of the band in queue order
                q[k] = run;                                      // where it starts
                q[JOB_BANDS + 1 + k] = run;                      // ... and where its launch's cursor starts
            }
            if (c == 0) q[JOB_BANDS] = run + hist[0];  // length of the queue
        }
    }
    __syncthreads();
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n_comp; c += gridDim.x * blockDim.x) {
        const int c0 = p.cgraph_ptr[c], n = p.cgraph_ptr[c + 1] - c0, g = p.cgraph_win[c];
        // not queued: empty, the graph was skipped (arc buffer overflow), or k_single_track has finished the component
        if (n <= 0 || p.n_paths[g] < 0 || p.first_sink[c] == FAST_DONE) continue;
        const int b = p.comp_class[c];
        p.job_tab[s_start[b] + atomicAdd(&cursor[b], 1)] = make_int4(c0, n, g, c);
    }
}

// ---- successive shortest paths on every component graph. Blocks fetch components from their band of the work queue.
//      Queue descriptor q: band k in queue order holds jobs [q[k], q[k + 1]) and is drained through cursor
//      q[JOB_BANDS + 1 + k]. The launches have no data dependency (disjoint jobs, results meet in integer atomics):
//      every launch lets the next one start right away (programmatic dependent launch) and waits for the one before it
//      when it runs out of work, so that whatever follows in the stream sees all of them finished.
template <int BT, int MINB, bool SHARED>
__global__ void __launch_bounds__(BT, MINB) k_ssp_solve(const __grid_constant__ SolveParams p) {
    __shared__ int s_scan[64];
    __shared__ KeyIdx s_red[33];
    __shared__ int s_ok[8], s_job;
    __shared__ int4 s_tab;
    __shared__ long long s_work[S_WORK_LEN];  // statistics of the component being solved
    const int tid = threadIdx.x;
    if (p.chain & 1) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int k = JOB_BANDS - 1 - p.role;
    const int32_t *const q = p.job_meta + JOB_Q;
    const int n_jobs = q[k + 1];
    int *const cursor = p.job_meta + JOB_Q + JOB_BANDS + 1 + k;
    const long long t_block0 = clock64();

    // The queue index of the next job is fetched while the current one is being solved.
    int next_job = 0;
    if (tid == 0) next_job = atomicAdd(cursor, 1);
    for (;;) {
        if (tid == 0) {
            s_job = next_job;
            if (next_job < n_jobs) s_tab = p.job_tab[next_job];
        }
        if (tid < S_WORK_LEN) s_work[tid] = 0;
        __syncthreads();
        const int job = s_job;
        const int4 jt = s_tab;
        __syncthreads();
        if (job >= n_jobs) {
            // [17] of the batch's first graph: cycles the blocks spent between their first and their last job
            if (tid == 0 && p.stats)
                atomicAdd((unsigned long long *)(p.stats + 17), (unsigned long long)(clock64() - t_block0));
            break;
        }
        if (tid == 0) next_job = atomicAdd(cursor, 1);
        const int c0 = jt.x, n = jt.y, g = jt.z, cg = jt.w;
        if (n <= 0) continue;  // empty, or the graph was skipped (arc buffer overflow)
        bool done = false;
        if (SHARED && n <= p.rcap && staging_bytes(n, 0) <= p.sbytes)
            done = solve_component_sm(p, cg, c0, n, g, s_scan, s_red, s_ok, s_work);
        if (!done) {
            // Beyond the shared-memory bands. Two solvers on staging areas in global memory: layer sweeps
            // (solve_global.cuh: a step per frame layer, cheap when few sweeps settle a branch - long thin components)
            // and bucketed rounds over pushed work lists (solve_big.cuh: a step per record of the longest new path,
            // indifferent to how often that path changes direction - crowded components, where the sweeps alternate
            // dozens of times). Crowded = at least BIG_MIN_WIDTH detections per frame layer; costs beyond 32 bits always
            // take the sweeps.
            bool sweeps = p.win_wide[g] != 0;
            if (!sweeps) {
                int starts = 0;
                for (int i = tid; i < n; i += BT)
                    starts += (i == 0 || p.tkey[p.comp_perm[c0 + i]] != p.tkey[p.comp_perm[c0 + i - 1]]) ? 1 : 0;
                const int layers = block_sum(starts, s_scan);
                sweeps = (long long)n < (long long)BIG_MIN_WIDTH * layers;
            }
            if (sweeps)
                solve_component_global(p, cg, c0, n, g, s_scan, s_red, s_ok, s_work, 0);
            else
                solve_component_big(p, cg, c0, n, g, s_scan, s_red, s_ok, s_work);
        }
    }
    if (p.chain & 2) asm volatile("griddepcontrol.wait;" ::: "memory");
}

// ---- the reference keeps the first path of a graph even when its cost is >= 0 (wrapper.cpp:205,220). Components
//      only keep paths that pay; a graph that ended with no path at all gets the cheapest "first path" among its
//      components (none of them was touched, so the path is a plain chain of unused detections in the first tree).
__global__ void k_first_path_rule(const SolveParams p) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= p.n_graphs) return;
    if (p.n_paths[g] != 0) return;
    const int c_begin = p.win_cbase[g];
    const int c_end = (g + 1 < p.n_graphs) ? p.win_cbase[g + 1] : p.n_cgraphs[0];
    int64_t best = INF;
    int q = -1;
    for (int c = c_begin; c < c_end; c++)
        if (p.first_sink[c] >= 0 && p.first_cost[c] < best) {
            best = p.first_cost[c];
            q = p.first_sink[c];
        }
    if (q < 0) return;  // no source->sink path at all: K stays 0 (build_graph returns None for such graphs)
    int nx = TERM;
    for (int step = 0; step < p.n_total; step++) {
        const int d = p.comp_perm[q];
        p.next_out[d] = nx;
        const int pq = p.ft[q].par;
        if (pq < 0) {
            p.prev_out[d] = TERM;
            break;
        }
        p.prev_out[d] = p.comp_perm[pq];
        nx = d;
        q = pq;
    }
    p.n_paths[g] = 1;
    p.total[g] = best;
    if (p.path_cost) p.path_cost[p.cgraph_ptr[c_begin]] = best;
}

// ---- accepted path costs per graph in the order the reference finds them (ascending: successive shortest paths never
//      get cheaper). The solver leaves them grouped by component inside the graph's slice, MOT3D_INF_COST elsewhere.
//      One block per graph: stable compaction of the entries into `scratch`, then every entry is written at its rank
//      (ties by position: equal values are indistinguishable), the rest of the slice is MOT3D_INF_COST.
//      Quadratic in the number of paths of ONE graph (tiles of the compacted list go through shared memory).
constexpr int PC_TILE = 2048;
__global__ void __launch_bounds__(256) k_order_path_costs(int n_graphs, const int32_t *__restrict__ graph_ptr,
                                                          int64_t *__restrict__ path_cost, int64_t *__restrict__ scratch,
                                                          int32_t *__restrict__ too_many, int max_paths) {
    __shared__ int s_scan[33];
    __shared__ int64_t s_tile[PC_TILE];
    const int g = blockIdx.x;
    const int g0 = graph_ptr[g];
    const int n = graph_ptr[g + 1] - g0;
    const int T = blockDim.x, tid = threadIdx.x;
    int64_t *v = path_cost + g0, *c = scratch + g0;
    int K = 0;
    for (int base = 0; base < n; base += T) {
        const int i = base + tid;
        const int64_t x = i < n ? v[i] : INF;
        const int f = x < INF_CUT ? 1 : 0;
        int tot;
        const int ex_ = block_excl_scan(f, s_scan, &tot);
        if (f) c[K + ex_] = x;
        K += tot;
    }
    __syncthreads();
    if (K > max_paths) {  // left grouped by component; the caller orders this graph some other way
        if (tid == 0) atomicOr(too_many, 1);
        return;
    }
    for (int i = tid; i < n; i += T) v[i] = INF;
    __syncthreads();
    for (int ib = 0; ib < K; ib += T) {
        const int i = ib + tid;
        const int64_t x = i < K ? c[i] : INF;
        int rank = 0;
        for (int jb = 0; jb < K; jb += PC_TILE) {
            const int m = min(PC_TILE, K - jb);
            __syncthreads();
            for (int j = tid; j < m; j += T) s_tile[j] = c[jb + j];
            __syncthreads();
            if (i < K)
                for (int j = 0; j < m; j++) {
                    const int64_t y = s_tile[j];
                    rank += (y < x || (y == x && jb + j < i)) ? 1 : 0;
                }
        }
        if (i < K) v[rank] = x;
    }
}

// Tracks of each graph, ordered by first detection (the reference enumerates S's successors in node order,
// mot3d/mot3d.py:175-176, and keeps the post-nodes of each path).
__global__ void k_extract_tracks(int n_graphs, const int32_t *__restrict__ graph_ptr, const int32_t *__restrict__ next,
                                 const int32_t *__restrict__ prev, int32_t *__restrict__ track_count,
                                 int32_t *__restrict__ track_ptr, int32_t *__restrict__ track_det) {
    __shared__ int s_scan[33];
    const int g = blockIdx.x;
    const int g0 = graph_ptr[g];
    const int n = graph_ptr[g + 1] - g0;
    const int T = blockDim.x, tid = threadIdx.x;
    int32_t *tp = track_ptr + g0 + g;
    int32_t *td = track_det + g0;
    // starts, in ascending detection order (kept in tp[] until replaced by offsets)
    int K = 0;
    for (int base = 0; base < n; base += T) {
        const int i = base + tid;
        const int f = (i < n && prev[g0 + i] == TERM) ? 1 : 0;
        int tot;
        const int ex_ = block_excl_scan(f, s_scan, &tot);
        if (f) tp[K + ex_] = i;
        K += tot;
    }
    __syncthreads();
    int run = 0;
    for (int base = 0; base < K; base += T) {
        const int t = base + tid;
        int start = -1, len = 0;
        if (t < K) {
            start = tp[t];
            for (int v = start + g0; v >= 0 && len < n; v = next[v]) len++;
        }
        int tot;
        const int off = run + block_excl_scan(len, s_scan, &tot);
        if (t < K) {
            tp[t] = off;
            int k = 0;
            for (int v = start + g0; v >= 0 && k < len; v = next[v]) td[off + k++] = v - g0;
        }
        run += tot;
    }
    if (tid == 0) {
        tp[K] = run;
        track_count[g] = K;
    }
}

static size_t a256(size_t b) { return (b + 255) & ~(size_t)255; }

// measurement hook (mot3d_solve_trace_events): events recorded at the kernel boundaries of the next solve of this thread
static thread_local cudaEvent_t *g_trace_events = nullptr;
static thread_local int g_trace_count = 0;
static inline void trace_mark(int k, cudaStream_t st) {
    if (g_trace_events && k < g_trace_count && g_trace_events[k]) cudaEventRecord(g_trace_events[k], st);
}

// ---- per-device facts and launch parameters are looked up once, not on every call
// Launch shapes of the bands: threads per block, blocks per SM. Band SM_BANDS = the global staging area.
#ifndef M3_B0T
#define M3_B0T 128
#define M3_B0B 10
#endif
#ifndef M3_B1T
#define M3_B1T 256
#define M3_B1B 3
#endif
static const int kBandThreads[SM_BANDS + 1] = {M3_B0T, M3_B1T, 512, 256};
static const int kBandBlocks[SM_BANDS + 1] = {M3_B0B, M3_B1B, 1, 3};
static const void *band_kernel(int b) {
    switch (b) {
        case 0: return (const void *)k_ssp_solve<M3_B0T, M3_B0B, true>;
        case 1: return (const void *)k_ssp_solve<M3_B1T, M3_B1B, true>;
        case 2: return (const void *)k_ssp_solve<512, 1, true>;
        default: return (const void *)k_ssp_solve<256, 3, false>;
    }
}
struct DevInfo {
    int dev, n_sm, smem_sm, smem_block;
    int band_bytes[SM_BANDS];   // dynamic shared memory of a block per band = the staging area of one component
    int band_blocks[SM_BANDS + 1];  // blocks per SM the hardware really grants
};
static std::mutex g_dev_mutex;
static DevInfo g_dev[MOT3D_MAX_DEVICES];
static bool g_dev_known[MOT3D_MAX_DEVICES];
static DevInfo *dev_info() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MOT3D_MAX_DEVICES) return nullptr;
    std::lock_guard<std::mutex> lock(g_dev_mutex);
    DevInfo *d = &g_dev[dev];
    if (g_dev_known[dev]) return d;
    memset(d, 0, sizeof(*d));
    d->dev = dev;
    if (cudaDeviceGetAttribute(&d->n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return nullptr;
    if (cudaDeviceGetAttribute(&d->smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev) != cudaSuccess) return nullptr;
    if (cudaDeviceGetAttribute(&d->smem_block, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) return nullptr;
    for (int b = 0; b <= SM_BANDS; b++) {
        cudaFuncAttributes fa;
        if (cudaFuncGetAttributes(&fa, band_kernel(b)) != cudaSuccess) return nullptr;
        int dyn = 0;
        if (b < SM_BANDS) {
            // a block's share of the SM minus the 1 KB the system reserves per block and the kernel's static part
            dyn = d->smem_sm / kBandBlocks[b] - 1024 - (int)fa.sharedSizeBytes;
            if (dyn + (int)fa.sharedSizeBytes > d->smem_block) dyn = d->smem_block - (int)fa.sharedSizeBytes;
            dyn &= ~15;
            if (dyn < 4096) return nullptr;
            d->band_bytes[b] = dyn;
            if (dyn > 48 * 1024 &&
                cudaFuncSetAttribute(band_kernel(b), cudaFuncAttributeMaxDynamicSharedMemorySize, dyn) != cudaSuccess)
                return nullptr;
        }
        int occ = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, band_kernel(b), kBandThreads[b], (size_t)dyn) != cudaSuccess)
            return nullptr;
        d->band_blocks[b] = occ > 0 ? occ : 1;
    }
    g_dev_known[dev] = true;
    return d;
}

}  // namespace m3

using namespace m3;

extern "C" {

size_t mot3d_solve_workspace_bytes(int64_t n_total, int32_t n_graphs, int64_t arc_capacity) {
    if (n_total < 0 || n_graphs < 0 || arc_capacity < 0) return 0;
    size_t n = (size_t)(n_total > 0 ? n_total : 1), g = (size_t)(n_graphs > 0 ? n_graphs : 1);
    size_t e = (size_t)(arc_capacity > 0 ? arc_capacity : 1);
    return a256(n * 8) * 3 + a256(n * 4) * 12 + a256(n * 16) + a256((n + g) * 4) * 2 + a256(g * 4) * 2 +
           a256(n * sizeof(FtTmp)) * 2 +
           a256(n * sizeof(GRec)) + a256((e + n + 8) * sizeof(StArc)) + a256(n * 16) * 3 + a256(n * 8) * 2 + a256((n / 16 + n + 8) * 4) +
           a256((e + 8) * 12) + 256 +
           a256(mot3d_components_workspace_bytes(n_total, n_graphs)) + 512 + a256(JOB_META_INTS * sizeof(int32_t));
}

int32_t mot3d_solve_batch(int32_t n_graphs, int64_t n_total, const int32_t *graph_ptr, int32_t max_graph_nodes,
                          const int32_t *tkey, int64_t arc_capacity, const int32_t *in_ptr, const int32_t *in_src,
                          const int64_t *in_cost, const int64_t *entry_cost, const int64_t *node_cost,
                          const int64_t *exit_cost, int32_t *next_out, int32_t *prev_out, int32_t *n_paths_out,
                          int64_t *total_cost_out, int64_t *path_cost_out, int64_t *stats_out, void *ws, size_t ws_bytes,
                          void *stream) {
    return m3::solve_batch_phased(SOLVE_ALL, 0, n_graphs, n_graphs, n_total, graph_ptr, max_graph_nodes, tkey,
                                  arc_capacity, in_ptr, in_src, in_cost, entry_cost, node_cost, exit_cost, next_out,
                                  prev_out, n_paths_out, total_cost_out, path_cost_out, stats_out, ws, ws_bytes, stream);
}

/* mot3d_solve_batch in phases (include/mot3d_b200.h) */
int32_t mot3d_solve_batch_phase(int32_t phases, int32_t g_first, int32_t g_count, int32_t n_graphs, int64_t n_total,
                                const int32_t *graph_ptr, int32_t max_graph_nodes, const int32_t *tkey,
                                int64_t arc_capacity, const int32_t *in_ptr, const int32_t *in_src, const int64_t *in_cost,
                                const int64_t *entry_cost, const int64_t *node_cost, const int64_t *exit_cost,
                                int32_t *next_out, int32_t *prev_out, int32_t *n_paths_out, int64_t *total_cost_out,
                                int64_t *path_cost_out, int64_t *stats_out, void *ws, size_t ws_bytes, void *stream) {
    M3_CHECK_ARG(phases > 0 && (phases & ~SOLVE_ALL) == 0, "phases: bits MOT3D_SOLVE_BEGIN / _FRONT / _BACK");
    return m3::solve_batch_phased(phases, g_first, g_count, n_graphs, n_total, graph_ptr, max_graph_nodes, tkey,
                                  arc_capacity, in_ptr, in_src, in_cost, entry_cost, node_cost, exit_cost, next_out,
                                  prev_out, n_paths_out, total_cost_out, path_cost_out, stats_out, ws, ws_bytes, stream);
}

}  // extern "C"

namespace m3 {
int32_t solve_batch_phased(int phases, int32_t g_first, int32_t g_count, int32_t n_graphs, int64_t n_total,
                           const int32_t *graph_ptr, int32_t max_graph_nodes, const int32_t *tkey, int64_t arc_capacity,
                           const int32_t *in_ptr, const int32_t *in_src, const int64_t *in_cost,
                           const int64_t *entry_cost, const int64_t *node_cost, const int64_t *exit_cost,
                           int32_t *next_out, int32_t *prev_out, int32_t *n_paths_out, int64_t *total_cost_out,
                           int64_t *path_cost_out, int64_t *stats_out, void *ws, size_t ws_bytes, void *stream) {
    M3_CHECK_ARG(n_graphs >= 0 && max_graph_nodes >= 0 && arc_capacity >= 0, "negative sizes");
    if (n_graphs == 0) return MOT3D_OK;
    M3_CHECK_ARG(graph_ptr && tkey && in_ptr && entry_cost && node_cost && exit_cost, "NULL input");
    M3_CHECK_ARG(next_out && prev_out && n_paths_out && total_cost_out && ws, "NULL output");
    M3_CHECK_ARG(n_total >= max_graph_nodes, "n_total < max_graph_nodes");
    if (ws_bytes < mot3d_solve_workspace_bytes(n_total, n_graphs, arc_capacity)) {
        set_error("solve workspace too small: %zu < %zu bytes", ws_bytes,
                  mot3d_solve_workspace_bytes(n_total, n_graphs, arc_capacity));
        return MOT3D_E_CAPACITY;
    }
    const size_t g = (size_t)n_graphs;
    const size_t n_cap = (size_t)(n_total > 0 ? n_total : 1);
    const size_t e_cap = (size_t)(arc_capacity > 0 ? arc_capacity : 1);
    char *w = (char *)ws;
    auto take = [&](size_t bytes) {
        char *r = w;
        w += a256(bytes);
        return r;
    };
    SolveParams p;
    p.n_graphs = n_graphs;
    p.g_first = 0;
    p.rcap = 0;
    p.sbytes = 0;
    p.chain = 0;
    p.role = 0;
    p.n_total = (int64_t)n_total;
    p.graph_ptr = graph_ptr;
    p.tkey = tkey;
    p.in_ptr = in_ptr;
    p.in_src = in_src;
    p.in_cost = in_cost;
    p.en = entry_cost;
    p.cn = node_cost;
    p.ex = exit_cost;
    p.next_out = next_out;
    p.prev_out = prev_out;
    p.n_paths = n_paths_out;
    p.total = total_cost_out;
    p.path_cost = path_cost_out;
    p.stats = stats_out;
    p.arc_capacity = arc_capacity;
    p.ft = (FtTmp *)take(n_cap * sizeof(FtTmp));
    p.first_cost = (int64_t *)take(n_cap * 8);
    p.ws_tk = (int32_t *)take(n_cap * 4);
    p.ws_bl = (int32_t *)take(n_cap * 4);
    p.ws_hop = (int32_t *)take(n_cap * 4);
    p.ws_pl = (int32_t *)take(n_cap * 4);
    p.first_sink = (int32_t *)take(n_cap * 4);
    p.job_tab = (int4 *)take(n_cap * 16);
    int32_t *comp_perm = (int32_t *)take(n_cap * 4);
    int32_t *comp_inv = (int32_t *)take(n_cap * 4);
    int32_t *cgraph_win = (int32_t *)take(n_cap * 4);
    p.ws_lvl = (int32_t *)take((n_cap + g) * 4);
    int32_t *cgraph_ptr = (int32_t *)take((n_cap + g) * 4);
    int32_t *win_cbase = (int32_t *)take(g * 4);
    p.win_wide = (int32_t *)take(g * 4);
    p.ws_rec = (GRec *)take(n_cap * sizeof(GRec));
    p.ws_arc = (StArc *)take((e_cap + n_cap + 8) * sizeof(StArc));
    const size_t cws_bytes = mot3d_components_workspace_bytes(n_total, n_graphs);
    void *cws = take(cws_bytes);
    int32_t *n_cgraphs = (int32_t *)take(16);
    p.job_meta = (int32_t *)take(JOB_META_INTS * sizeof(int32_t));
    p.comp_class = (int32_t *)take(n_cap * 4);
    p.ft_tmp = (FtTmp *)take(n_cap * sizeof(FtTmp));
    p.ft_anc = (int32_t *)take(n_cap * 4);
    p.big_res_e = (longlong2 *)take(n_cap * 16);
    p.big_res_f = (int4 *)take(n_cap * 16);
    p.big_listed = (unsigned *)take((n_cap / 16 + n_cap + 8) * 4);
    p.big_old_e = (longlong2 *)take(n_cap * 16);
    p.big_prio = (long long *)take(n_cap * 8);
    p.big_far = (int *)take(n_cap * 8);
    p.big_delta = BIG_DELTA;
    double ov;
    if (option(OPT_BIG_DELTA, &ov) && ov >= 0) p.big_delta = (long long)ov;
    p.big_arc_stride = (int64_t)e_cap + 8;
    p.big_arc = (int32_t *)take((e_cap + 8) * 12);
    p.big_arc_next = (unsigned long long *)take(256);
    p.comp_perm = comp_perm;
    p.comp_inv = comp_inv;
    p.cgraph_ptr = cgraph_ptr;
    p.cgraph_win = cgraph_win;
    p.win_cbase = win_cbase;
    p.n_cgraphs = n_cgraphs;

    cudaStream_t st = (cudaStream_t)stream;
    const DevInfo *di = dev_info();
    if (!di) return cuda_fail(cudaGetLastError(), "device attributes / shared memory limits of the solver kernels");
    M3_CHECK_ARG(g_first >= 0 && g_count >= 0 && g_first + g_count <= n_graphs, "graph range outside the batch");
    if (phases & SOLVE_BEGIN) {
        M3_CUDA(cudaMemsetAsync(p.job_meta, 0, JOB_META_INTS * sizeof(int32_t), st));
        M3_CUDA(cudaMemsetAsync(p.big_arc_next, 0, sizeof(unsigned long long), st));
        M3_CUDA(cudaMemsetAsync(p.first_sink, 0xff, n_cap * sizeof(int32_t), st));  // -1 = component has no first path
    }
    trace_mark(0, st);
    // 1. first shortest-path tree per graph: 2 lanes per node, 256 threads: a layer of ~100 detections is one pass of
    //    the block
#ifndef M3_FT_T
#define M3_FT_T 256
#endif
    if ((phases & SOLVE_FRONT) && g_count > 0) {
        p.g_first = g_first;
        k_first_tree<2><<<g_count, M3_FT_T, 0, st>>>(p);
        M3_LAUNCH_CHECK("k_first_tree");
    }
    trace_mark(1, st);
    // 2. connected components of every graph (from the branches of the tree and the arcs between them), detections
    //    and tree regrouped by component
    FtOut tree;
    tree.tmp = p.ft_tmp;
    tree.anc = p.ft_anc;
    tree.out = p.ft;
    int32_t rc = components_run(n_graphs, n_total, graph_ptr, max_graph_nodes, in_ptr, in_src, arc_capacity, comp_perm,
                                comp_inv, cgraph_ptr, cgraph_win, win_cbase, n_cgraphs, cws, cws_bytes, stream, g_first,
                                (phases & SOLVE_FRONT) ? g_count : 0, (phases & SOLVE_BACK) != 0, &tree);
    if (rc) return rc;
    trace_mark(2, st);
    if (!(phases & SOLVE_BACK)) return MOT3D_OK;
    // 3. successive shortest paths per component, one launch per band of the work queue (see k_job_order)
    const int n_sm = di->n_sm;
    for (int b = 0; b < SM_BANDS; b++) {
        p.band_bytes[b] = di->band_bytes[b];
        p.band_rec[b] = SM_MAX_RECORDS;
    }
    // test hook (tests/test_gpu_parity.py::test_single_track_pass_equals_full_solver): every component through the
    // full solver
    p.fast = (option(OPT_SOLVE_NO_FAST, &ov) && ov != 0) ? 0 : 1;
    {
        // one warp per component, 8 x 8 warps per SM striding over them: finished on the spot (single tracks) or
        // counted into its band and size class; then the queue is laid out
        long long blocks = (long long)n_sm * 8;
        if (blocks * 8 > n_total) blocks = (n_total + 7) / 8;
        if (blocks < 1) blocks = 1;
        k_single_track<<<(unsigned)blocks, 256, 0, st>>>(p);
        M3_LAUNCH_CHECK("k_single_track");
        long long jblocks = (n_total + 256 * 8 - 1) / (256 * 8);
        if (jblocks > n_sm) jblocks = n_sm;
        if (jblocks < 1) jblocks = 1;
        k_job_order<<<(unsigned)jblocks, 256, 0, st>>>(p);
    }
    M3_LAUNCH_CHECK("k_job_order");
    trace_mark(3, st);
    // Which shared-memory bands can hold a component of this batch at all: a component of n detections needs at most
    // staging_bytes(n, n (n - 1) / 2). The launch for the global staging area is always there (costs beyond 32 bits
    // are only known on the device); so is band 0.
    const long long mg = max_graph_nodes;
    int hi_band = SM_BANDS - 1;
    for (int b = SM_BANDS - 1; b >= 0; b--)
        if (mg <= p.band_rec[b] && staging_bytes(mg, mg * (mg - 1) / 2) <= di->band_bytes[b]) hi_band = b;
    bool chained = false;  // a launch before this one lets it start early (programmatic dependent launch)
    for (int b = SM_BANDS; b >= 0; b--) {
        if (b < SM_BANDS && b > hi_band) continue;
        SolveParams pb = p;
        pb.role = b;
        pb.chain = (b > 0 ? 1 : 0) | (chained ? 2 : 0);
        pb.rcap = b < SM_BANDS ? p.band_rec[b] : 0;
        pb.sbytes = b < SM_BANDS ? di->band_bytes[b] : 0;
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = chained ? 1 : 0;
        chained = true;
        long long grid = (long long)n_sm * di->band_blocks[b];
        // never more blocks than components the band can hold (only band 0 takes components of fewer than 32 detections)
        const long long most = b == 0 ? n_total : n_total / 32 + 1;
        if (grid > most) grid = most > 0 ? most : 1;
        cfg.gridDim = dim3((unsigned)grid);
        cfg.blockDim = dim3((unsigned)kBandThreads[b]);
        cfg.dynamicSmemBytes = (size_t)pb.sbytes;
        switch (b) {
            case 0: M3_CUDA(cudaLaunchKernelEx(&cfg, k_ssp_solve<M3_B0T, M3_B0B, true>, pb)); break;
            case 1: M3_CUDA(cudaLaunchKernelEx(&cfg, k_ssp_solve<M3_B1T, M3_B1B, true>, pb)); break;
            case 2: M3_CUDA(cudaLaunchKernelEx(&cfg, k_ssp_solve<512, 1, true>, pb)); break;
            default: M3_CUDA(cudaLaunchKernelEx(&cfg, k_ssp_solve<256, 3, false>, pb)); break;
        }
    }
    M3_LAUNCH_CHECK("k_ssp_solve");
    trace_mark(4, st);
    // 4. the reference's "first path is always kept" rule, per graph
    k_first_path_rule<<<(n_graphs + 127) / 128, 128, 0, st>>>(p);
    M3_LAUNCH_CHECK("k_first_path_rule");
    trace_mark(5, st);
    return MOT3D_OK;
}
}  // namespace m3

extern "C" {

int32_t mot3d_order_path_costs(int32_t n_graphs, const int32_t *graph_ptr, int64_t *path_cost, int64_t *scratch,
                               int32_t max_paths, int32_t *too_many_flag, void *stream) {
    M3_CHECK_ARG(n_graphs >= 0 && max_paths >= 0, "negative sizes");
    if (n_graphs == 0) return MOT3D_OK;
    M3_CHECK_ARG(graph_ptr && path_cost && scratch && too_many_flag, "NULL pointer");
    k_order_path_costs<<<n_graphs, 256, 0, (cudaStream_t)stream>>>(n_graphs, graph_ptr, path_cost, scratch, too_many_flag,
                                                                  max_paths);
    M3_LAUNCH_CHECK("k_order_path_costs");
    return MOT3D_OK;
}

int32_t mot3d_solve_trace_events(void **events, int32_t n_events) {
    g_trace_events = (cudaEvent_t *)events;
    g_trace_count = events ? n_events : 0;
    return MOT3D_OK;
}

int32_t mot3d_extract_tracks(int32_t n_graphs, const int32_t *graph_ptr, const int32_t *next, const int32_t *prev,
                             int32_t *track_count_out, int32_t *track_ptr_out, int32_t *track_det_out, void *stream) {
    M3_CHECK_ARG(n_graphs >= 0, "n_graphs < 0");
    if (n_graphs == 0) return MOT3D_OK;
    M3_CHECK_ARG(graph_ptr && next && prev && track_count_out && track_ptr_out && track_det_out, "NULL pointer");
    k_extract_tracks<<<n_graphs, 256, 0, (cudaStream_t)stream>>>(n_graphs, graph_ptr, next, prev, track_count_out,
                                                                track_ptr_out, track_det_out);
    M3_LAUNCH_CHECK("k_extract_tracks");
    return MOT3D_OK;
}

}  // extern "C"
