// Min-cost-flow solve: successive shortest paths on the residual network, one thread block per connected component.
//
// Replaces wrapper.solve (reference mot3d/solvers/wrappers/muSSP/wrapper.cpp:171-317) and the muSSP core it drives
// (zip!/muSSP-master/muSSP/Graph.cpp).  Same network, same path-update rule, same acceptance rule, exact int64:
//   network        S -> pre_i (entry), pre_i -> post_i (node cost), post_i -> T (exit), post_i -> pre_j (transition);
//                  unit capacities                                   (mot3d/mot3d.py:75-141)
//   first tree     one relaxation pass over the frame layers         (Graph::shortest_path_dag, Graph.cpp:139-187)
//   components     a graph falls apart into connected components (csrc/components.cu); each one is a job of the work
//                  queue, staged ONCE into the block's shared memory (records = its detections, arc lists = their
//                  in-arcs; every label -- parent, branch, next, prev -- is a slot of that staging area), solved path
//                  after path without global traffic, and only its flow is written back. Components that do not fit
//                  (or whose costs need more than 32 bits) run the same code on a staging area in global memory.
//   next paths     muSSP's observation (Graph.cpp:341-343, 394-503): after a path through S -> pre_a is flipped,
//                  only the branch of the shortest-path tree that hung from that arc loses its labels; every
//                  other node keeps its exact distance.  The branch is listed in frame-layer order and re-labelled
//                  by an exact Bellman-Ford fix-point in pull mode: ascending layer order settles the forward arcs,
//                  a descending pass settles the reversed arcs (they only run backwards in time along tracks); the
//                  two alternate until nothing moves.  No tolerances, no heap.
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

constexpr int LPN = 4;  // lanes cooperating on one node's in-arcs (first tree, initial labels)

// state markers (slots / positions are >= 0)
constexpr int NONE = -1;  // detection unused
constexpr int TERM = -2;  // next: track ends here (post->T saturated); prev: track starts here (S->pre saturated)
constexpr int PAR_S = -1;     // pre reached from S
constexpr int PAR_POST = -2;  // pre reached from its own post (reversed node arc)
constexpr int PAR_NONE = -3;
constexpr int DL_PRE = 1 << 29;  // branch flags of a record: its pre / post node is being re-labelled
constexpr int DL_POST = 1 << 30;
constexpr int DL_FIRST = 1 << 28;  // first record of its frame layer in the branch list

// ---- staging area of a component, two flavours:
//   SM = true   shared memory; costs in 32 bits (k_first_tree checked that they fit; NO_ARC32 = no such arc), slots
//               in 16 bits. 64-byte records, 8-byte arcs.
//   SM = false  global workspace; 64-bit costs, 32-bit slots: any size, any cost. 80-byte records, 16-byte arcs.
// A record, in 16-byte groups so that a relaxation is a handful of 128-bit loads:
//   A  x = DL_* flags of the current branch, y = arc list (see arcs_of), z = prev, w = next (slots / NONE / TERM)
//   F  x = parent of the pre node (slot / PAR_*), y = branch of pre, z = branch of post (slot of the detection whose
//      S -> pre arc starts the node's tree path; muSSP's ancestor_node_id, Graph.cpp:160-161)
//   E  exact distances from S of pre and post (the invariant the incremental update relies on)
//   W  costs: S -> pre, pre -> post, post -> T, matched arc prev -> this
constexpr int NO_ARC32 = 0x7fffffff;
__device__ __forceinline__ bool fits_narrow(int64_t v) {
    return v >= INF_CUT || (v == (int64_t)(int)v && (int)v != NO_ARC32);
}
struct __align__(16) Cost4 {
    int64_t x, y, z, w;
};
struct __align__(8) PArc {  // the first arc of a list also carries the length of the list in pad
    int cost;
    unsigned short src;
    unsigned short pad;
};
struct __align__(16) StArc {
    long long cost;
    int src;
    int pad;
};
template <bool SM>
struct Staging;
template <>
struct Staging<true> {
    typedef int4 costs_t;
    typedef int cost_t;
    typedef PArc arc_t;
    typedef unsigned short slot_t;
    static __device__ __forceinline__ int conv(int64_t v) { return v >= INF_CUT ? NO_ARC32 : (int)v; }
    static __device__ __forceinline__ bool has(int v) { return v != NO_ARC32; }
};
template <>
struct Staging<false> {
    typedef Cost4 costs_t;
    typedef int64_t cost_t;
    typedef StArc arc_t;
    typedef int slot_t;
    static __device__ __forceinline__ int64_t conv(int64_t v) { return v; }
    static __device__ __forceinline__ bool has(int64_t v) { return v < INF_CUT; }
};
template <bool SM>
struct __align__(16) RecT {
    int4 A, F;
    longlong2 E;
    typename Staging<SM>::costs_t W;
};
static_assert(sizeof(RecT<true>) == 64 && sizeof(RecT<false>) == 80, "record layout");
constexpr int SM_BYTES_PER_REC = sizeof(RecT<true>) + 4 + 3 * 2;  // + layer key + entries of the three slot lists
constexpr int SM_ARC_BYTES = sizeof(PArc);
// shared-memory footprint of a component of n detections with `arcs` in-arcs (every list also carries its length)
__host__ __device__ __forceinline__ long long staging_bytes(long long n, long long arcs) {
    const long long n8 = (n + 7) & ~7ll;
    return n8 * SM_BYTES_PER_REC + 16 * ((n8 >> 5) + 3) + (arcs + n + 8) * SM_ARC_BYTES;
}

struct SolveParams {
    int32_t n_graphs;
    int32_t rcap;    // shared-memory staging area of this launch: most records a component may have ...
    int32_t sbytes;  // ... and its size in bytes (records, arc lists and slot lists of one component, laid out per component)
    int32_t role;   // 0 = the whole queue; 1 / 2 / 3 = the launch of band 2 / 1 / 0 (largest components first)
    // queue bands: a component goes to the first band whose staging area holds it (records and bytes, see
    // staging_bytes); band 2 also takes what fits nowhere (solved on the global staging area)
    int32_t band_rec[2];
    int32_t band_bytes[2];
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
    // first shortest-path tree (k_first_tree), per regrouped position; values are regrouped positions too
    int64_t *ws_dpre;
    int64_t *ws_dpost;
    int32_t *ws_par;       // parent of the pre node: PAR_S / PAR_NONE / position of the source detection
    int32_t *ws_anc_pre;   // branch = position of the detection whose S->pre arc starts the node's tree path
    int32_t *ws_anc_post;
    int32_t *ws_lvl;       // [n_total + n_graphs] frame layers
    // global staging area (components that do not fit shared memory; spilled arc lists of those that do)
    RecT<false> *ws_rec;   // [n_total]
    int32_t *ws_tk;        // [n_total] layer keys
    int32_t *ws_bl;        // [n_total] x 3 slot lists
    int32_t *ws_hop;
    int32_t *ws_pl;
    StArc *ws_arc;         // [arc_capacity + n_total + 8]: the list of detection d starts at in_ptr[d] + d
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

extern __shared__ __align__(16) unsigned char smem_raw[];

// Work counters and phase clocks of the component a block is solving live in shared memory (s_work[10 ..]) and are
// kept by thread 0 only: they are statistics, they must not cost every thread registers.
//   s_work[10 + k], k < 7  cycles: arg-min, branch flags / list, staging + initial labels, walk, fix-point, -, certificate
//   s_work[17 ..]          arcs read while staging, branches re-labelled, records listed for them, components finished by
//                          the optimality certificate instead of a last re-labelling
struct Counters {
    long long *w;
    long long t_mark, t_start;
    __device__ __forceinline__ void tick(int k) {
        if (threadIdx.x == 0) {
            const long long now = clock64();
            w[10 + k] += now - t_mark;
            t_mark = now;
        }
    }
    __device__ __forceinline__ void add(int k, long long v) {
        if (threadIdx.x == 0) w[17 + k] += v;
    }
};
constexpr int S_WORK_LEN = 21;

// ---- first shortest-path tree of every graph (before it is split into components): no flow yet, the network is a
//      DAG, one pull sweep over the frame layers with the whole block (Graph::shortest_path_dag, Graph.cpp:139-187).
//      State is written in the regrouped order; branch label = position of the detection whose S arc starts the path.
constexpr int FT_W = 2048;  // ring of post / branch labels (power of two)
template <int FT_LPN>
__global__ void __launch_bounds__(256, 5) k_first_tree(const SolveParams p) {
    __shared__ int s_scan[33];
    __shared__ int64_t s_dp[FT_W];
    __shared__ int s_an[FT_W];
    const int g = blockIdx.x;
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
    const int32_t *inv = p.comp_inv + g0;
    if ((int64_t)in_ptr[n] > p.arc_capacity) {  // arc buffer overflow (reported by the fill kernel): skipped
        if (tid == 0) p.n_paths[g] = -1;
        for (int i = tid; i < n; i += T) {
            p.next_out[g0 + i] = NONE;
            p.prev_out[g0 + i] = NONE;
        }
        return;
    }
    for (int i = tid; i < n; i += T) {  // regrouped positions of this graph are [g0, g0 + n) as well
        const int q = g0 + i;  // (distances, parent and branch labels are written by the sweep, reachable or not)
        p.first_cost[q] = INF;
        p.first_sink[q] = -1;
        if (p.path_cost) p.path_cost[q] = INF;
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
            int q = 0;
            int64_t e_ = INF, c_ = 0, x_ = INF;
            int64_t w_best = INF, w_min = INF;  // cost of the parent arc / of the cheapest in-arc (k_single_track)
            if (valid) {
                const int e0 = in_ptr[j], deg = in_ptr[j + 1] - e0;
                if (sub == 0) {  // the node's own costs: in flight together with the arc loads
                    q = inv[j];
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
                            di[u] = p.ws_dpost[p.comp_inv[src[u]]];
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
                    anc = best.idx >= ring_lo ? s_an[(best.idx - g0) & (FT_W - 1)] : p.ws_anc_post[p.comp_inv[best.idx]];
                wide |= (fits_narrow(e_) && fits_narrow(x_) && c_ == (int64_t)(int)c_) ? 0 : 1;
                if (e_ < best.key) {
                    best.key = e_;
                    best.idx = PAR_S;
                    anc = q;
                }
                // what solve_component's optimality certificate asks of this detection if its tree arc ends up
                // matched: no negative entry / exit cost; the arc is the cheapest in-arc, costs mc <= 0, cn + mc <= 0
                int cert = ((e_ < INF_CUT && e_ < 0) || (x_ < INF_CUT && x_ < 0)) ? 0 : 1;
                if (best.idx >= 0 && !(w_best <= 0 && c_ + w_best <= 0 && w_min >= w_best)) cert = 0;
                p.ws_pl[q] = cert;
                int64_t dpost = INF;
                if (best.key < INF_CUT) {
                    dpost = best.key + c_;
                } else {  // unreachable
                    best.key = INF;
                    best.idx = PAR_NONE;
                    anc = -1;
                }
                p.ws_dpre[q] = best.key;
                p.ws_par[q] = best.idx;  // a detection index for now
                p.ws_anc_pre[q] = anc;
                p.ws_dpost[q] = dpost;
                p.ws_anc_post[q] = anc;
                s_dp[j & (FT_W - 1)] = dpost;
                s_an[j & (FT_W - 1)] = anc;
            }
        }
        __syncthreads();
    }
    for (int i = tid; i < n; i += T) {  // parents: detection index -> regrouped position
        const int par = p.ws_par[g0 + i];
        if (par >= 0) p.ws_par[g0 + i] = p.comp_inv[par];
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

// Solve one component: positions [c0, c0 + n) of graph g. SM = staging area in shared memory (see Staging).
// RPT = records per thread in the relabelling rounds of the shared-memory flavour (staging capacity <= RPT x block).
template <bool SM, int JAC_RPT>
__device__ __noinline__ void solve_component(const SolveParams &p, const int cg, const int c0, const int n, const int g,
                                             int *s_scan, KeyIdx *s_red, int *s_ok, long long *s_work,
                                             const int sweep_warp) {
    typedef Staging<SM> S;
    typedef RecT<SM> Rec;
    typedef typename S::arc_t Arc;
    typedef typename S::costs_t Costs;
    typedef typename S::cost_t Cost;
    typedef typename S::slot_t Slot;
    const int T = blockDim.x, tid = threadIdx.x;
    const int sub = tid % LPN, grp = tid / LPN, ngrp = T / LPN;
    const unsigned gmask = 0xFu << ((tid & 31) / LPN * LPN);
    const int n8 = (n + 7) & ~7;
    const int racap = SM ? (p.sbytes - n8 * SM_BYTES_PER_REC - 16 * ((n8 >> 5) + 3)) / SM_ARC_BYTES - 8 : 0;  // arcs the shared area holds
    // slot lists: bl = the branch in layer order, hop = detections whose matched in-arc changed, pl = the branch's
    // unused records with a post node to re-label ("producers" of the ascending sweep)
    Rec *rec;
    Arc *arc;
    int32_t *tk;
    Slot *bl, *hop, *pl;
    if (SM) {
        // laid out for this component: records, slot lists, then every byte that is left for the arc lists
        rec = (Rec *)smem_raw;
        tk = (int32_t *)(smem_raw + (size_t)n8 * sizeof(Rec));
        bl = (Slot *)(tk + n8);
        hop = bl + n8;
        pl = hop + n8;
        arc = (Arc *)(smem_raw + (size_t)n8 * SM_BYTES_PER_REC + 16 * ((n8 >> 5) + 3));  // (after the change bits)
    } else {
        rec = (Rec *)(p.ws_rec + c0);
        arc = nullptr;
        tk = p.ws_tk + c0;
        bl = (Slot *)(p.ws_bl + c0);
        hop = (Slot *)(p.ws_hop + c0);
        pl = (Slot *)(p.ws_pl + c0);
    }
    // arc list of a record: A.y >= 0 = offset in the shared arc area, < 0 = ~offset in the global area. The global
    // area is addressed in StArc units whatever the arc type, so that lists of different detections never overlap.
    auto arcs_of = [&](int ay) -> Arc * { return (!SM || ay < 0) ? (Arc *)(p.ws_arc + ~ay) : arc + ay; };
    const int32_t *perm = p.comp_perm + c0;

    Counters cnt;
    cnt.w = s_work;
    cnt.t_mark = cnt.t_start = clock64();

    // ---- stage the component. Records first: every load of a record is issued before any is used (one memory level
    //      after the permutation); its in-degree and first arc wait in A.x / F.w for the arc pass. Arc-list offsets by
    //      a block scan over contiguous chunks.
    const int chunk = (n + T - 1) / T;
    const int i0 = min(tid * chunk, n), i1 = min(i0 + chunk, n);
    if (SM && tid == 0) s_ok[2] = 0;  // largest layer span of an arc (SM flavour, see the relabelling rounds)
    {
        int my_arcs = 0, my_layers = 0, prev_key = 0;
        bool first_flag = false;
        if (SM && i0 < i1 && i0 > 0) prev_key = p.tkey[perm[i0 - 1]];
        for (int i = i0; i < i1; i++) {
            const int q = c0 + i, d = perm[i];
            const int e0 = p.in_ptr[d], e1 = p.in_ptr[d + 1];
            const int par = p.ws_par[q], ap_ = p.ws_anc_pre[q], ao_ = p.ws_anc_post[q];
            const int64_t dpre = p.ws_dpre[q], dpost = p.ws_dpost[q];
            const int64_t en = p.en[d], cn = p.cn[d], ex = p.ex[d];
            const int key = p.tkey[d];
            Rec r;
            r.A = make_int4(e1 - e0, 0, NONE, NONE);
            r.F = make_int4(par >= 0 ? par - c0 : par, ap_ >= 0 ? ap_ - c0 : -1, ao_ >= 0 ? ao_ - c0 : -1, e0);
            r.E = make_longlong2(dpre, dpost);
            r.W.x = S::conv(en);
            r.W.y = S::conv(cn);
            r.W.z = S::conv(ex);
            r.W.w = 0;
            rec[i] = r;
            tk[i] = key;
            my_arcs += e1 - e0 + 1;  // + 1: a list without arcs still carries its length
            if (SM) {  // a new frame layer starts here
                const bool f = i > 0 && key != prev_key;
                if (i == i0) first_flag = f;
                my_layers += f ? 1 : 0;
                prev_key = key;
            }
        }
        int tot;
        int a0 = block_excl_scan1(my_arcs, s_scan, 0, &tot);
        int l0 = 0;
        if (SM) {
            int nl;
            l0 = block_excl_scan1(my_layers, s_scan, 1, &nl);
        }
        for (int i = i0; i < i1; i++) {
            const int deg = rec[i].A.x;
            int off = a0;
            a0 += deg + 1;
            // (shared arc area full:) the list lives in global memory, at the detection's own place
            if (!SM || off + deg + 1 > racap) off = ~(rec[i].F.w + perm[i]);
            rec[i].A.y = off;
            if (SM) {  // the SM flavour keeps the rank of the record's frame layer instead of its time key
                const int key = tk[i];
                if (i == i0 ? first_flag : key != prev_key) l0++;
                prev_key = key;
                tk[i] = l0;
            }
        }
        cnt.add(0, tot - n);
    }
    __syncthreads();
    int span = 0;
    for (int ib = 0; ib < n; ib += ngrp) {  // arc lists: one lane group per record, two independent arcs per lane
        const int i = ib + grp;
        if (i < n) {
            const int e0 = rec[i].F.w, deg = rec[i].A.x;
            Arc *const ap = arcs_of(rec[i].A.y);
            const int my_layer = SM ? tk[i] : 0;
            for (int kb = 0; kb < deg; kb += 2 * LPN) {
                const int k0 = kb + sub, k1 = kb + LPN + sub;
                const int src0 = k0 < deg ? p.in_src[e0 + k0] : 0, src1 = k1 < deg ? p.in_src[e0 + k1] : 0;
                const int64_t w0 = k0 < deg ? p.in_cost[e0 + k0] : 0, w1 = k1 < deg ? p.in_cost[e0 + k1] : 0;
                const int q0 = p.comp_inv[src0], q1 = p.comp_inv[src1];
                if (k0 < deg) {
                    Arc x;
                    x.cost = (decltype(x.cost))w0;
                    x.src = (decltype(x.src))(q0 - c0);
                    x.pad = (decltype(x.pad))deg;  // only read from the first arc of the list
                    ap[k0] = x;
                    if (SM) span = max(span, my_layer - tk[q0 - c0]);
                }
                if (k1 < deg) {
                    Arc x;
                    x.cost = (decltype(x.cost))w1;
                    x.src = (decltype(x.src))(q1 - c0);
                    x.pad = (decltype(x.pad))deg;
                    ap[k1] = x;
                    if (SM) span = max(span, my_layer - tk[q1 - c0]);
                }
            }
            if (deg == 0 && sub == 0) {
                Arc x;
                x.cost = 0;
                x.src = 0;
                x.pad = 0;
                ap[0] = x;
            }
        }
    }
    if (SM) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) span = max(span, __shfl_xor_sync(0xffffffffu, span, d));
        if (lane_id() == 0 && span > 0) atomicMax(&s_ok[2], span);
    }
    __syncthreads();
    cnt.tick(2);

    int K = 0, n_used = 0;
    int64_t total = 0;
    for (int iter = 0; iter <= n; iter++) {  // at most one path per detection
        // ---- cheapest way into T
        KeyIdx best;
        best.key = INF;
        best.idx = -1;
        for (int i = tid; i < n; i += T) {
            const int64_t d = rec[i].E.y;
            const Cost x = rec[i].W.z;
            if (rec[i].A.w != TERM && d < INF_CUT && S::has(x)) {
                KeyIdx v;
                v.key = d + x;
                v.idx = i;
                best = min_ki(best, v);
            }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const KeyIdx o = shfl_xor_ki(best, d);
            const bool take = o.key < best.key || (o.key == best.key && (unsigned)o.idx < (unsigned)best.idx);
            best.key = take ? o.key : best.key;
            best.idx = take ? o.idx : best.idx;
        }
        if (lane_id() == 0) s_red[warp_id()] = best;
        __syncthreads();
        KeyIdx sink;
        sink.key = INF;
        sink.idx = -1;
        for (int k = 0; k < (T + 31) >> 5; k++) {
            const KeyIdx o = s_red[k];
            if (o.key < sink.key || (o.key == sink.key && (unsigned)o.idx < (unsigned)sink.idx)) sink = o;
        }
        __syncthreads();  // s_red is rewritten next iteration
        cnt.tick(0);
        if (sink.idx < 0) break;  // T unreachable
        if (sink.key >= 0) {      // see k_ssp_solve: the first-path rule belongs to the whole graph
            if (K == 0 && tid == 0) {
                p.first_cost[cg] = sink.key;
                p.first_sink[cg] = c0 + sink.idx;
            }
            break;
        }
        const int root = rec[sink.idx].F.z;
        if (root < 0) break;  // cannot happen: a labelled node always has a branch
        // ---- augment: walk the path back from T and rewrite next/prev (Graph::flip_path, Graph.cpp:266-335)
        if (tid == 0) {
            int ok = 0, nhop = 0, nxt_new = TERM, cur = sink.idx, joined = 0;
            for (int step = 0; step <= n; step++) {
                const int4 A = rec[cur].A;         // z = prev, w = next
                const int par_cur = rec[cur].F.x;  // in flight together: the parent of pre_cur (used when cur joins)
                rec[cur].A.w = nxt_new;
                int m, q;
                if (A.z == NONE) {  // post_cur came from pre_cur: detection cur joins a track
                    m = cur;
                    q = par_cur;
                    joined++;
                } else {  // post_cur came from pre_m, m = its old successor: the arc cur -> m is cancelled
                    m = A.w;
                    q = PAR_NONE;
                    while (m >= 0) {
                        q = rec[m].F.x;
                        if (q != PAR_POST) break;
                        const int m2 = rec[m].A.w;  // pre_m came from post_m: m leaves its track
                        rec[m].A.w = NONE;
                        rec[m].A.z = NONE;
                        m = m2;
                        joined--;
                    }
                }
                if (m < 0) break;  // cannot happen on a consistent residual state
                if (q < 0) {  // PAR_S: the path starts here
                    rec[m].A.z = TERM;
                    ok = 1;
                    break;
                }
                rec[m].A.z = q;
                hop[nhop++] = (Slot)m;  // the cost of the newly matched arc q -> m is looked up below
                nxt_new = m;
                cur = q;
            }
            s_ok[0] = ok;
            s_ok[1] = nhop;
            s_ok[7] = joined;
        }
        __syncthreads();
        cnt.tick(3);
        if (!s_ok[0]) break;  // defensive: inconsistent state, stop with what we have
        n_used += s_ok[7];
        for (int h = tid; h < s_ok[1]; h += T) {  // cost of every newly matched arc prev -> m
            const int m = hop[h], q = rec[m].A.z;
            const Arc *const ap = arcs_of(rec[m].A.y);
            const int deg = ap[0].pad;
            for (int k = 0; k < deg; k++) {
                const Arc x = ap[k];
                if ((int)x.src == q) {
                    rec[m].W.w = x.cost;
                    break;
                }
            }
        }
        __syncthreads();
        // ---- Is this component finished? Once every detection is on a track, a further S -> T path in the residual
        //      network has the shape  S -> pre_i, (pre -> post of the predecessor: cancels a matched arc, -mc), then
        //      any number of hops post_a -> pre_b -> post_prev(b) (an unused arc and the cancellation of b's matched
        //      arc: w_ab - mc_b) or post_a -> pre_a -> post_prev(a) (-cn_a - mc_a), then post -> T. If every matched arc
        //      is the cheapest in-arc of its head (w_ab >= mc_b), mc <= 0, cn + mc <= 0 and no entry / exit cost is
        //      negative, every piece is >= 0: no path can pay, exactly what the re-labelling + arg-min below would
        //      find out. One pass over the arcs instead.
        if (n_used == n) {
            bool good = true;
            for (int ib = 0; ib < n; ib += ngrp) {
                const int i = ib + grp;
                if (i < n) {
                    const int4 A = rec[i].A;
                    const Costs W = rec[i].W;
                    if ((S::has(W.x) && W.x < 0) || (S::has(W.z) && W.z < 0)) good = false;
                    if (A.z >= 0) {
                        const int64_t mc = W.w;
                        if (mc > 0 || (int64_t)W.y + mc > 0) good = false;
                        const Arc *const ap = arcs_of(A.y);
                        const int deg = ap[0].pad;
                        for (int k = sub; k < deg; k += LPN) {
                            const Arc x = ap[k];
                            if ((int)x.src != A.z && (int64_t)x.cost < mc) good = false;
                        }
                    }
                }
            }
            if (__syncthreads_and(good)) {
                if (tid == 0 && p.path_cost) p.path_cost[c0 + K] = sink.key;
                K++;
                total += sink.key;
                cnt.add(3, 1);
                cnt.tick(6);
                break;
            }
            cnt.tick(6);
        }
        if constexpr (SM) {
            // ---- re-label the branch hanging from S -> pre_root: synchronous Bellman-Ford rounds on the staged
            //      component, ONE THREAD PER RECORD. Every node of the branch starts unlabelled; a round computes,
            //      from the labels of the round before, the best label of the record's pre node (in-arcs, S arc,
            //      reversed node arc) and post node (own pre node when unused, reversed matched arc of the successor
            //      when on a track); all threads then publish at once. Labels only fall, the labels outside the
            //      branch are exact and never move (Graph.cpp:341-343), so the fix-point is the exact distance of
            //      every node; it is reached after as many rounds as the new shortest paths have records, whatever
            //      their shape - forward arcs and reversed arcs travel in the same round, there are no alternating
            //      sweeps and no sequential layer steps. A parent is only replaced by a strictly better one, read
            //      from the round before: the parent graph stays a tree (a cycle of tight arcs would need every arc
            //      set strictly later than the one before it) and the result does not depend on timing.
            if (tid == 0 && p.path_cost) p.path_cost[c0 + K] = sink.key;
            //      A record is only looked at in a round when one of its inputs may have moved in the round before: a
            //      post label in one of the `span` frame layers below it, or - for a record on a track - a pre label in
            //      one of the `span` layers above it (its successor lives there). One bit per layer and kind, two
            //      buffers each (read / being written): a superset of what an exact dependency list would give, for
            //      no adjacency memory and a handful of register operations per record and round.
            //      The records to look at are compacted into a work list (pl), so that a round with a handful of them
            //      costs one warp a single pass through the relaxation code: four lanes per record when they fit,
            //      one lane per record otherwise.
            const int span = s_ok[2];
            const int nwords = (n8 >> 5) + 3;
            unsigned *const chg = (unsigned *)(pl + n8);  // [post | pre][2][nwords]
            int ent[JAC_RPT], lay[JAC_RPT], nxt[JAC_RPT];
            int n_mine = 0;
#pragma unroll
            for (int k = 0; k < JAC_RPT; k++) {
                const int s = tid + k * T;
                ent[k] = 0;
                lay[k] = 0;
                nxt[k] = NONE;
                if (s < n) {
                    int4 F = rec[s].F;
                    const int f = (F.y == root ? DL_PRE : 0) | (F.z == root ? DL_POST : 0);
                    ent[k] = f;
                    if (f) {
                        longlong2 E = rec[s].E;
                        if (f & DL_PRE) {
                            E.x = INF;
                            F.x = PAR_NONE;
                            F.y = -1;
                        }
                        if (f & DL_POST) {
                            E.y = INF;
                            F.z = -1;
                        }
                        rec[s].E = E;
                        rec[s].F = F;
                        lay[k] = tk[s];
                        nxt[k] = rec[s].A.w;
                        n_mine++;
                    }
                }
            }
            for (int w = tid; w < 4 * nwords; w += T) chg[w] = 0;
            if (tid == 0) s_ok[3] = 0;  // length of the round's work list
            cnt.add(1, 1);
            __syncthreads();
            cnt.tick(1);
            int n_arc_pull = 0, n_relax = 0, rounds = 0;
            const long long t_fix0 = clock64();
            const unsigned lt = (1u << lane_id()) - 1u;
            auto window = [&](const unsigned *bits, int lo, int width) -> bool {  // any of bits [lo, lo + width)?
                const unsigned long long two = (unsigned long long)bits[lo >> 5] | ((unsigned long long)bits[(lo >> 5) + 1] << 32);
                return width > 0 && ((two >> (lo & 31)) & ((1ull << width) - 1ull)) != 0ull;
            };
            // New labels of record s from the labels of the round before; LPR lanes share its in-arcs (sub = my share).
            // Returns 1 if the pre label moved, + 2 if the post label moved.
            auto relax = [&](auto lpr_tag, const int s, const bool do_pre, const bool do_post, const int sub,
                             int64_t &dpre, int64_t &dpost, int &par, int &apre, int &apost) -> int {
                constexpr int LPR = decltype(lpr_tag)::value;
                const int4 A = rec[s].A;  // z = prev, w = next
                const Costs W = rec[s].W;
                const longlong2 E = rec[s].E;
                const int4 F = rec[s].F;
                dpre = E.x;
                dpost = E.y;
                par = F.x;
                apre = F.y;
                apost = F.z;
                if (do_post && A.w >= 0) {  // on a track: reversed matched arc pre_next -> post_s
                    const int64_t dm = rec[A.w].E.x;
                    if (dm < INF_CUT) {
                        const int64_t cand = dm - (int64_t)rec[A.w].W.w;
                        if (cand < dpost) {
                            dpost = cand;
                            apost = rec[A.w].F.y;
                        }
                    }
                }
                if (do_pre) {
                    const Arc *const ap = arcs_of(A.y);
                    const int deg = ap[0].pad;
                    int64_t key = INF;
                    int kpar = 0x7fffffff;
                    if (LPR == 1) {
                        for (int kb = 0; kb < deg; kb += 4) {  // four independent arc / label loads in flight
                            Arc x[4];
                            int64_t d[4];
#pragma unroll
                            for (int u = 0; u < 4; u++) x[u] = ap[kb + u < deg ? kb + u : 0];
#pragma unroll
                            for (int u = 0; u < 4; u++) d[u] = rec[x[u].src].E.y;
#pragma unroll
                            for (int u = 0; u < 4; u++) {
                                const int64_t cand = d[u] + (int64_t)x[u].cost;
                                const bool ok = kb + u < deg && (int)x[u].src != A.z && d[u] < INF_CUT;
                                if (ok && (cand < key || (cand == key && (int)x[u].src < kpar))) {
                                    key = cand;  // lowest slot wins ties
                                    kpar = x[u].src;
                                }
                            }
                        }
                    } else {
                        for (int kb = sub; kb < deg; kb += 2 * LPR) {  // two independent arc / label loads in flight
                            const bool two = kb + LPR < deg;
                            const Arc x0 = ap[kb], x1 = ap[two ? kb + LPR : kb];
                            const int64_t d0 = rec[x0.src].E.y, d1 = rec[x1.src].E.y;
                            const int64_t c0_ = d0 + (int64_t)x0.cost, c1_ = d1 + (int64_t)x1.cost;
                            if ((int)x0.src != A.z && d0 < INF_CUT && (c0_ < key || (c0_ == key && (int)x0.src < kpar))) {
                                key = c0_;
                                kpar = x0.src;
                            }
                            if (two && (int)x1.src != A.z && d1 < INF_CUT && (c1_ < key || (c1_ == key && (int)x1.src < kpar))) {
                                key = c1_;
                                kpar = x1.src;
                            }
                        }
#pragma unroll
                        for (int d = 1; d < LPR; d <<= 1) {
                            const int64_t ok_ = __shfl_xor_sync(gmask, key, d);
                            const int op = __shfl_xor_sync(gmask, kpar, d);
                            if (ok_ < key || (ok_ == key && op < kpar)) {
                                key = ok_;
                                kpar = op;
                            }
                        }
                    }
                    if (sub == 0) n_arc_pull += deg;
                    int kanc = -1;
                    if (key < INF_CUT)
                        kanc = rec[kpar].F.z;
                    else
                        kpar = PAR_NONE;
                    if (A.z != TERM && S::has(W.x) && (int64_t)W.x < key) {
                        key = W.x;
                        kpar = PAR_S;
                        kanc = s;
                    }
                    if (A.z != NONE && dpost < INF_CUT && dpost - (int64_t)W.y < key) {  // reversed node arc
                        key = dpost - (int64_t)W.y;
                        kpar = PAR_POST;
                        kanc = apost;
                    }
                    if (key < dpre) {
                        dpre = key;
                        par = kpar;
                        apre = kanc;
                    }
                }
                if (do_post && A.z == NONE && dpre < INF_CUT && dpre + (int64_t)W.y < dpost) {
                    dpost = dpre + (int64_t)W.y;  // unused record: post hangs from the own pre node
                    apost = apre;
                }
                return (dpre != E.x ? 1 : 0) | (dpost != E.y ? 2 : 0);
            };
            for (; rounds < 4 * n + 8; rounds++) {
                const long long t_r0 = clock64();
                const int rd = rounds & 1;
                const unsigned *const post_rd = chg + rd * nwords, *const pre_rd = chg + (2 + rd) * nwords;
                unsigned *const post_wr = chg + (rd ^ 1) * nwords, *const pre_wr = chg + (2 + (rd ^ 1)) * nwords;
#pragma unroll
                for (int k = 0; k < JAC_RPT; k++) {  // (A) which of my records may have an input that moved?
                    bool dirty = false;
                    if (ent[k]) {
                        dirty = rounds == 0 || span > 32;
                        if (!dirty && (ent[k] & DL_PRE)) {
                            const int lo = max(lay[k] - span, 0);
                            dirty = window(post_rd, lo, lay[k] - lo);
                        }
                        if (!dirty && (ent[k] & DL_POST) && nxt[k] >= 0) dirty = window(pre_rd, lay[k] + 1, span);
                    }
                    const unsigned m = __ballot_sync(0xffffffffu, dirty);
                    if (m) {
                        int base = 0;
                        if (lane_id() == 0) base = atomicAdd(&s_ok[3], __popc(m));
                        base = __shfl_sync(0xffffffffu, base, 0);
                        if (dirty)
                            pl[base + __popc(m & lt)] = (Slot)((tid + k * T) | ((ent[k] & DL_PRE) ? 0x4000 : 0) |
                                                               ((ent[k] & DL_POST) ? 0x8000 : 0));
                    }
                }
                __syncthreads();
                const int n_work = s_ok[3];
                long long t_a = 0;
                if (tid == 0) { t_a = clock64(); s_work[5] += t_a - t_r0; }
                if (n_work == 0) break;  // nothing moved in the round before: fix-point
                int64_t n_pre[JAC_RPT], n_post[JAC_RPT];
                int n_par[JAC_RPT], n_apre[JAC_RPT], n_apost[JAC_RPT], w_slot[JAC_RPT];
                int upd[JAC_RPT];  // 1 = pre label moved, 2 = post label moved
#pragma unroll
                for (int k = 0; k < JAC_RPT; k++) upd[k] = 0;
                if (n_work <= ngrp) {  // (B) relax them: four lanes per record ...
                    if (grp < n_work) {
                        const int we = pl[grp];
                        w_slot[0] = we & 0x3fff;
                        upd[0] = relax(std::integral_constant<int, LPN>(), w_slot[0], (we & 0x4000) != 0, (we & 0x8000) != 0,
                                       sub, n_pre[0], n_post[0], n_par[0], n_apre[0], n_apost[0]);
                        if (sub != 0) upd[0] = 0;
                    }
                    if (tid == 0) n_relax += n_work;
                } else {  // ... or one thread per record
#pragma unroll
                    for (int k = 0; k < JAC_RPT; k++) {
                        const int wi = tid + k * T;
                        if (wi >= n_work) continue;
                        const int we = pl[wi];
                        w_slot[k] = we & 0x3fff;
                        n_relax++;
                        upd[k] = relax(std::integral_constant<int, 1>(), w_slot[k], (we & 0x4000) != 0, (we & 0x8000) != 0, 0,
                                       n_pre[k], n_post[k], n_par[k], n_apre[k], n_apost[k]);
                    }
                }
                __syncthreads();  // every label of the round before has been read
                long long t_b = 0;
                if (tid == 0) { t_b = clock64(); s_work[7] += t_b - t_a; if (n_work <= ngrp) s_work[1] += 1; }
                for (int w = tid; w < nwords; w += T) {  // the next round's write buffers
                    chg[rd * nwords + w] = 0;
                    chg[(2 + rd) * nwords + w] = 0;
                }
                if (tid == 0) s_ok[3] = 0;
#pragma unroll
                for (int k = 0; k < JAC_RPT; k++)  // (C) publish
                    if (upd[k]) {
                        const int s = w_slot[k];
                        rec[s].E = make_longlong2(n_pre[k], n_post[k]);
                        int4 F = rec[s].F;
                        F.x = n_par[k];
                        F.y = n_apre[k];
                        F.z = n_apost[k];
                        rec[s].F = F;
                        const int l = tk[s];
                        if (upd[k] & 1) atomicOr(&pre_wr[l >> 5], 1u << (l & 31));
                        if (upd[k] & 2) atomicOr(&post_wr[l >> 5], 1u << (l & 31));
                    }
                __syncthreads();
                if (tid == 0) s_work[9] += clock64() - t_b;
            }
            // work counters of this branch
            {
                int nodes = n_relax;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) {
                    n_arc_pull += __shfl_xor_sync(0xffffffffu, n_arc_pull, d);
                    nodes += __shfl_xor_sync(0xffffffffu, nodes, d);
                    n_mine += __shfl_xor_sync(0xffffffffu, n_mine, d);
                }
                if (lane_id() == 0) {
                    atomicAdd((unsigned long long *)&s_work[2], (unsigned long long)(2 * nodes));
                    atomicAdd((unsigned long long *)&s_work[3], (unsigned long long)n_arc_pull);
                    atomicAdd((unsigned long long *)&s_work[8], (unsigned long long)n_mine);
                }
                if (tid == 0) {
                    s_work[0] += rounds + 1;
                    s_work[6] += rounds + 1;
                    s_work[4] += clock64() - t_fix0;
                }
            }
        } else {
            // ---- the branch hanging from S -> pre_root, as a list of slots in layer order
            int nd;
            {
                int my_nd = 0;
                for (int i = i0; i < i1; i++) {
                    const int4 F = rec[i].F;
                    const int f = (F.y == root ? DL_PRE : 0) | (F.z == root ? DL_POST : 0);
                    rec[i].A.x = f;
                    my_nd += f ? 1 : 0;
                }
                int lp = block_excl_scan1(my_nd, s_scan, 1, &nd);
                if (my_nd)
                    for (int i = i0; i < i1; i++)
                        if (rec[i].A.x) bl[lp++] = (Slot)i;
            }
            __syncthreads();
            for (int lp = tid; lp < nd; lp += T)
                if (lp == 0 || tk[bl[lp]] != tk[bl[lp - 1]]) rec[bl[lp]].A.x |= DL_FIRST;
            if (tid == 0 && p.path_cost) p.path_cost[c0 + K] = sink.key;
            cnt.add(1, 1);
            cnt.add(2, nd);
            cnt.tick(1);
            {  // the "producers" of the ascending sweeps (see below): every thread looks at a chunk of the list
                const int lchunk = (nd + T - 1) / T;
                const int lp0 = min(tid * lchunk, nd), lp1 = min(lp0 + lchunk, nd);
                int mine = 0;
                for (int lp = lp0; lp < lp1; lp++) {
                    const int4 A = rec[bl[lp]].A;
                    mine += ((A.x & DL_POST) && A.z == NONE) ? 1 : 0;
                }
                int np;
                int k = block_excl_scan1(mine, s_scan, 0, &np);
                if (mine)
                    for (int lp = lp0; lp < lp1; lp++) {
                        const int4 A = rec[bl[lp]].A;
                        if ((A.x & DL_POST) && A.z == NONE) pl[k++] = bl[lp];
                    }
                if (tid == 0) {
                    s_ok[6] = np;
                    s_ok[2] = s_ok[3] = s_ok[4] = s_ok[5] = 0;  // flags of the fix-point rounds
                }
            }
            __syncthreads();
            // ---- new labels of the branch start from what cannot change during the fix-point: in-arcs from outside the
            //      branch, the S arc, the own post node when it is outside. One lane group per record.
            for (int sb = 0; sb < nd; sb += ngrp) {
                const int lp = sb + grp;
                const bool valid = lp < nd;
                const int s = valid ? bl[lp] : 0;
                const int4 A = rec[s].A;
                KeyIdx cb;
                cb.key = INF;
                cb.idx = -1;
                if (valid && (A.x & DL_PRE)) {
                    const Arc *const ap = arcs_of(A.y);
                    const int deg = ap[0].pad;
                    for (int k = sub; k < deg; k += LPN) {
                        const Arc x = ap[k];
                        const int64_t d = rec[x.src].E.y;
                        if (!(rec[x.src].A.x & DL_POST) && (int)x.src != A.z && d < INF_CUT) {
                            const int64_t cand = d + x.cost;
                            if (cand < cb.key || (cand == cb.key && (unsigned)x.src < (unsigned)cb.idx)) {
                                cb.key = cand;
                                cb.idx = x.src;
                            }
                        }
                    }
                }
    #pragma unroll
                for (int d = 1; d < LPN; d <<= 1) {
                    const KeyIdx o = shfl_xor_ki(cb, d, gmask);
                    const bool take = o.key < cb.key || (o.key == cb.key && (unsigned)o.idx < (unsigned)cb.idx);
                    cb.key = take ? o.key : cb.key;
                    cb.idx = take ? o.idx : cb.idx;
                }
                // every lane of every group has finished reading its sources' post labels before any of them is reset
                // (a record with DL_POST is never a source here, but its E/F words are rewritten as a whole)
                __syncwarp();
                if (valid && sub == 0) {
                    int4 F = rec[s].F;
                    const longlong2 E = rec[s].E;
                    const Costs W = rec[s].W;
                    int64_t key = INF, dpost_new = E.y;
                    int par = PAR_NONE, anc = -1;
                    if (A.x & DL_PRE) {
                        if (cb.idx >= 0) {
                            key = cb.key;
                            par = cb.idx;
                            anc = rec[cb.idx].F.z;
                        }
                        if (!(A.x & DL_POST) && A.z != NONE && E.y < INF_CUT && E.y - W.y < key) {
                            key = E.y - W.y;  // own post node is outside the branch: a constant candidate
                            par = PAR_POST;
                            anc = F.z;
                        }
                        if (A.z != TERM && S::has(W.x) && W.x < key) {
                            key = W.x;
                            par = PAR_S;
                            anc = s;
                        }
                        if (key >= INF_CUT) {
                            key = INF;
                            par = PAR_NONE;
                            anc = -1;
                        }
                        F.x = par;
                        F.y = anc;
                    }
                    if (A.x & DL_POST) {
                        const bool unused_labelled = (A.x & DL_PRE) && A.z == NONE && key < INF_CUT;
                        dpost_new = unused_labelled ? key + W.y : INF;
                        F.z = unused_labelled ? anc : -1;
                    }
                    rec[s].E = make_longlong2((A.x & DL_PRE) ? key : E.x, dpost_new);
                    rec[s].F = F;
                }
            }
            __syncthreads();
            cnt.tick(2);
            // ---- fix-point (see relabel_branch). Ascending sweep in two phases: (A) the unused records whose post label
            //      is recomputed from their pre label are the only ones that pass a new label on to higher layers, they
            //      go first, layer by layer, by one warp; (B) every other record only reads post labels that are final
            //      for this sweep, so all of them are relaxed at once by the whole block, 8 lanes per record.
            //      Descending pass: independent chain walks, one thread per track segment.
            {
                const int lane = lane_id(), wid = warp_id(), nw = (T + 31) >> 5;
                const int r8 = lane >> 3, a8 = lane & 7;
                int n_node32 = 0, n_arc32 = 0, n_astep = 0, n_dstep = 0;
                long long t_asc = 0, t_desc = 0, t0 = clock64();
                const int np = s_ok[6];  // producers (list pl)
                // relax the pre node of record s over its in-arcs, S and its own post node, then its post node; 8 lanes.
                // Returns true if a change can travel along a reversed arc (to lower layers).
                auto relax = [&](const int s, const bool valid) -> bool {
                    const Rec *rp = rec + s;
                    const int4 A = rp->A;
                    const Costs W = rp->W;
                    int4 F = rp->F;
                    const longlong2 E = rp->E;
                    const int ent = valid ? A.x : 0;
                    KeyIdx best;
                    best.key = INF;
                    best.idx = -1;
                    int deg = 0;
                    if (ent & DL_PRE) {
                        const Arc *const ap = arcs_of(A.y);
                        const Arc x0 = ap[a8];  // speculative: the first arc also carries the list length
                        deg = ap[0].pad;
                        if (a8 < deg) {
                            const int64_t d = rec[x0.src].E.y;
                            if ((int)x0.src != A.z && d < INF_CUT) {
                                best.key = d + x0.cost;
                                best.idx = x0.src;
                            }
                        }
                        for (int k = a8 + 8; k < deg; k += 8) {
                            const Arc x = ap[k];
                            const int64_t d = rec[x.src].E.y;
                            if ((int)x.src != A.z && d < INF_CUT &&
                                (d + x.cost < best.key || (d + x.cost == best.key && (unsigned)x.src < (unsigned)best.idx))) {
                                best.key = d + x.cost;
                                best.idx = x.src;
                            }
                        }
                    }
    #pragma unroll
                    for (int d = 1; d < 8; d <<= 1) {  // lowest slot wins ties
                        const KeyIdx o = shfl_xor_ki(best, d);
                        const bool take = o.key < best.key || (o.key == best.key && (unsigned)o.idx < (unsigned)best.idx);
                        best.key = take ? o.key : best.key;
                        best.idx = take ? o.idx : best.idx;
                    }
                    bool down = false;
                    if (ent != 0 && a8 == 0) {
                        int64_t dpre_s = E.x, dpost_s = E.y;
                        bool upd = false;
                        if (ent & DL_PRE) {
                            int64_t key = INF;
                            int par = PAR_NONE, anc = -1;
                            if (best.idx >= 0) {
                                key = best.key;
                                par = best.idx;
                                anc = rec[best.idx].F.z;
                            }
                            if (A.z != TERM && S::has(W.x) && W.x < key) {
                                key = W.x;
                                par = PAR_S;
                                anc = s;
                            }
                            if (A.z != NONE && dpost_s < INF_CUT && dpost_s - W.y < key) {
                                key = dpost_s - W.y;
                                par = PAR_POST;
                                anc = F.z;
                            }
                            if (key < dpre_s) {
                                dpre_s = key;
                                F.x = par;
                                F.y = anc;
                                upd = true;
                                down = A.z >= 0;  // pre_s feeds the post node of its predecessor on the track
                            }
                            n_arc32 += deg;
                        }
                        // The post node of a record on a track hangs from its successor's pre node: that is the descending
                        // pass's business, and it keeps every post label read during phase (B) stable.
                        if ((ent & DL_POST) && A.z == NONE && dpre_s < INF_CUT && dpre_s + W.y < dpost_s) {
                            dpost_s = dpre_s + W.y;
                            F.z = F.y;
                            upd = true;
                        }
                        if (upd) {
                            rec[s].E = make_longlong2(dpre_s, dpost_s);
                            rec[s].F = F;
                        }
                    }
                    return down;
                };
                for (int round = 0; round < 2 * nd + 4; round++) {
                    const int pb = round & 1;
                    if (round > 0) {  // round 0: the initial labels stand in for the ascending sweep
                        bool down = false;
                        if (wid == sweep_warp) {  // (A) producers, one layer at a time
                            for (int q = 0; q < np;) {
                                const int k = q + r8;
                                const bool in_range = k < np;
                                const int s = pl[in_range ? k : np - 1];
                                const bool same = in_range && tk[s] == tk[pl[q]];
                                const unsigned m = __ballot_sync(0xffffffffu, a8 == 0 && !same) & 0x01010100u;
                                const int cn_t = m ? (__ffs(m) - 1) >> 3 : 4;
                                down |= relax(s, r8 < cn_t);
                                n_node32 += cn_t;
                                n_astep++;
                                q += cn_t;
                                __syncwarp();
                            }
                        }
                        __syncthreads();
                        for (int base = 0; base < nd; base += 4 * nw) {  // (B) everything else, all at once
                            const int lp = base + 4 * wid + r8;
                            const bool in_range = lp < nd;
                            const int s = bl[in_range ? lp : nd - 1];
                            const int f = rec[s].A.x;
                            const bool producer = (f & DL_POST) && rec[s].A.z == NONE;
                            down |= relax(s, in_range && !producer);
                            n_astep++;
                        }
                        if (wid == 0) n_node32 += nd - np;
                        if (__any_sync(0xffffffffu, down) && lane == 0) s_ok[2 + pb] = 1;
                        if (wid == sweep_warp && lane == 0) s_work[0] += 1;
                        __syncthreads();
                        {
                            const long long t1 = clock64();
                            t_asc += t1 - t0;
                            t0 = t1;
                        }
                        if (!s_ok[2 + pb]) break;  // nothing can travel down: fix-point
                    }
                    // descending pass: every thread looks at its share of the list
                    bool up = false;
                    for (int lp = tid; lp < nd; lp += T) {
                        const int s0 = bl[lp];
                        int4 A = rec[s0].A;
                        if (A.z == NONE) {
                            if (A.x & DL_POST) {
                                const longlong2 E = rec[s0].E;
                                const Cost cn = rec[s0].W.y;
                                if (E.x < INF_CUT && E.x + cn < E.y) {
                                    rec[s0].E.y = E.x + cn;
                                    rec[s0].F.z = rec[s0].F.y;
                                    up = true;
                                }
                            }
                            continue;
                        }
                        if ((A.x & DL_POST) && A.w >= 0) continue;  // reached from its successor's walk
                        // on a track, successor not in the branch: a segment starts here; walk it down
                        int cur = s0;
                        Costs W = rec[cur].W;
                        int4 F = rec[cur].F;
                        longlong2 E = rec[cur].E;
                        int64_t nx_dpre = INF, nx_mc = 0;
                        int nx_apre = -1;
                        for (;;) {
                            // the predecessor's post node hangs from pre_cur: it is in the branch when pre_cur is
                            const int prv = ((A.x & DL_PRE) && A.z >= 0) ? A.z : -1;
                            int4 An = A, Fn = F;
                            Costs Wn = W;
                            longlong2 En = E;
                            if (prv >= 0) {  // fetch the next record of the walk before working on this one
                                const Rec *rn = rec + prv;
                                An = rn->A;
                                Wn = rn->W;
                                Fn = rn->F;
                                En = rn->E;
                            }
                            const int ent = A.x;
                            int64_t dpre_s = E.x, dpost_s = E.y;
                            bool upd = false;
                            if ((ent & DL_POST) && nx_dpre < INF_CUT && nx_dpre - nx_mc < dpost_s) {
                                dpost_s = nx_dpre - nx_mc;
                                F.z = nx_apre;
                                upd = true;
                                up = true;  // forward arcs leave post_s: the ascending sweep has to come back
                            }
                            if ((ent & DL_PRE) && (ent & DL_POST) && dpost_s < INF_CUT && dpost_s - W.y < dpre_s) {
                                dpre_s = dpost_s - W.y;
                                F.x = PAR_POST;
                                F.y = F.z;
                                upd = true;
                            }
                            if (upd) {
                                rec[cur].E = make_longlong2(dpre_s, dpost_s);
                                rec[cur].F = F;
                            }
                            n_dstep++;
                            if (prv < 0) break;
                            nx_dpre = dpre_s;
                            nx_mc = W.w;
                            nx_apre = F.y;
                            cur = prv;
                            A = An;
                            W = Wn;
                            F = Fn;
                            E = En;
                        }
                    }
                    if (__any_sync(0xffffffffu, up) && lane == 0) s_ok[4 + pb] = 1;
                    if (tid == 0) {
                        s_ok[2 + (pb ^ 1)] = 0;  // the next round's flags
                        s_ok[4 + (pb ^ 1)] = 0;
                        s_work[1] += 1;
                    }
                    __syncthreads();
                    {
                        const long long t1 = clock64();
                        t_desc += t1 - t0;
                        t0 = t1;
                    }
                    if (!s_ok[4 + pb] && round > 0) break;  // nothing can travel up: fix-point
                }
                // work counters of this branch
    #pragma unroll
                for (int d = 16; d > 0; d >>= 1) {
                    n_arc32 += __shfl_xor_sync(0xffffffffu, n_arc32, d);
                    n_dstep += __shfl_xor_sync(0xffffffffu, n_dstep, d);
                }
                if (lane == 0) {
                    atomicAdd((unsigned long long *)&s_work[2], (unsigned long long)(n_node32 + n_dstep));
                    atomicAdd((unsigned long long *)&s_work[3], (unsigned long long)n_arc32);
                    atomicAdd((unsigned long long *)&s_work[7], (unsigned long long)n_dstep);
                    if (wid == sweep_warp) {
                        s_work[4] += t_asc;
                        s_work[5] += t_desc;
                        atomicAdd((unsigned long long *)&s_work[6], (unsigned long long)n_astep);
                    }
                }
            }
        }
        __syncthreads();
        cnt.tick(4);
        K++;
        total += sink.key;
    }
    // ---- results of this component: flow in detection indices, totals added to its graph
    __syncthreads();
    if (tid == 0) {
        if (K) {
            atomicAdd(&p.n_paths[g], K);
            atomicAdd((unsigned long long *)&p.total[g], (unsigned long long)total);
        }
        if (p.stats) {
            unsigned long long *o = (unsigned long long *)(p.stats + (size_t)MOT3D_SOLVE_STATS * g);
            atomicAdd(&o[0], (unsigned long long)s_work[0]);
            atomicAdd(&o[1], (unsigned long long)s_work[1]);
            atomicAdd(&o[2], (unsigned long long)s_work[2]);
            atomicAdd(&o[3], (unsigned long long)(s_work[17] + s_work[3]));
            for (int k = 0; k < 6; k++) atomicAdd(&o[4 + k], (unsigned long long)s_work[10 + k]);
            atomicAdd(&o[9], (unsigned long long)s_work[16]);
            atomicAdd(&o[10], (unsigned long long)s_work[20]);
            atomicAdd(&o[11], (unsigned long long)s_work[18]);
            atomicAdd(&o[12], (unsigned long long)(s_work[19] + s_work[8]));
            for (int k = 0; k < 4; k++) atomicAdd(&o[13 + k], (unsigned long long)s_work[4 + k]);
            if (SM) { atomicAdd(&o[18], (unsigned long long)s_work[5]); atomicAdd(&o[19], (unsigned long long)s_work[7]); atomicAdd(&o[20], (unsigned long long)s_work[9]); atomicAdd(&o[21], (unsigned long long)s_work[1]); }
            if (!SM) {  // the same four for the components on the global staging area alone, and how many they were
                for (int k = 0; k < 4; k++) atomicAdd(&o[18 + k], (unsigned long long)s_work[4 + k]);
                atomicAdd(&o[22], 1ull);
                atomicAdd(&o[23], (unsigned long long)(clock64() - cnt.t_start));
            }
        }
    }
    for (int i = tid; i < n; i += T) {
        const int a = rec[i].A.w, b = rec[i].A.z;
        const int d = perm[i];
        p.next_out[d] = a >= 0 ? perm[a] : a;
        p.prev_out[d] = b >= 0 ? perm[b] : b;
    }
    __syncthreads();
}

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
    return 4 * (lg - 1) + ((n >> (lg - 2)) & 3);  // 4 -> 4, 8 -> 8, ..., 8192 -> 48
}
constexpr int FAST_DONE = -2;
constexpr int JOB_BANDS = 3;
constexpr int JOB_Q = 2 * JOB_BANDS * JOB_CLASSES;  // offset of the queue descriptor in job_meta
constexpr int JOB_META_INTS = JOB_Q + 8;
__device__ __forceinline__ int band_class(const SolveParams &p, int n, long long arcs) {
    const long long need = staging_bytes(n, arcs);
    const int band = (n <= p.band_rec[0] && need <= p.band_bytes[0]) ? 0
                     : (n <= p.band_rec[1] && need <= p.band_bytes[1]) ? 1 : 2;
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
        const int b = band_class(p, n, arcs);
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
        const int par = p.ws_par[q];
        ok = ((i == 0) ? (par == PAR_S) : (par == q - 1)) && p.ws_pl[q] != 0;
        const int64_t ex = p.ex[p.comp_perm[q]];
        const int64_t dpost = p.ws_dpost[q];
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
    // three bands of size classes, the band of the largest components first: every band is drained by a launch of its
    // own whose staging area fits its components (see mot3d_solve_batch). The class counts were taken by
    // k_single_track; every block turns them into class starts and scatters its share of the components.
    constexpr int NB = JOB_BANDS * JOB_CLASSES;
    __shared__ int s_start[NB];
    const int n_comp = p.n_cgraphs[0];
    const int32_t *hist = p.job_meta;
    int32_t *cursor = p.job_meta + NB;
    if (threadIdx.x < NB) {
        int run = 0;
        for (int b = NB - 1; b > (int)threadIdx.x; b--) run += hist[b];
        s_start[threadIdx.x] = run;
        if (blockIdx.x == 0) {
            int32_t *q = p.job_meta + JOB_Q;
            if (threadIdx.x == 2 * JOB_CLASSES - 1) {  // everything above = band 2
                q[1] = run;
                q[4] = run;  // cursor of the launch for band 1 starts where band 2 ends
            }
            if (threadIdx.x == JOB_CLASSES - 1) {
                q[2] = run;
                q[5] = run;
            }
            if (threadIdx.x == 0) q[0] = run + hist[0];  // length of the queue
        }
    }
    __syncthreads();
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n_comp; c += gridDim.x * blockDim.x) {
        const int c0 = p.cgraph_ptr[c], n = p.cgraph_ptr[c + 1] - c0, g = p.cgraph_win[c];
        // not queued: empty, the graph was skipped (arc buffer overflow), or k_single_track has finished the component
        if (n <= 0 || p.n_paths[g] < 0 || p.first_sink[c] == FAST_DONE) continue;
        const int b = p.comp_class[c];
        p.job_tab[s_start[b] + atomicAdd(&cursor[b], 1)] = make_int4(c0, n, g, p.win_wide[g] ? (c | (1 << 31)) : c);
    }
}

// ---- successive shortest paths on every component graph. Blocks fetch components from a work queue.
template <int BT, int MINB>
__global__ void __launch_bounds__(BT, MINB) k_ssp_solve(const __grid_constant__ SolveParams p) {
    __shared__ int s_scan[64];
    __shared__ KeyIdx s_red[33];
    __shared__ int s_ok[8], s_sweep_warp, s_job;
    __shared__ int4 s_tab;
    // counters of the fix-point rounds: ascending sweeps, descending passes, node pulls, arc pulls, cycles in ascending
    // sweeps / descending passes, ascending steps / descending hops
    __shared__ long long s_work[S_WORK_LEN];

    const int T = blockDim.x;
    const int tid = threadIdx.x;
    // Which warp runs the sequential part of the ascending sweeps: blocks sharing an SM should use different SM
    // sub-partitions (warp schedulers), otherwise those warps fight for one issue port. %warpid is the hardware warp
    // slot; slots of a block are consecutive, slot % 4 = sub-partition. Pick the warp whose sub-partition equals the
    // block's slot group.
    if (tid % 32 == 0) {
        unsigned hw;
        asm volatile("mov.u32 %0, %%warpid;" : "=r"(hw));
        if (tid == 0) s_sweep_warp = 0;
        s_scan[warp_id()] = (int)hw;
    }
    __syncthreads();
    if (tid == 0) {
        const int nw = (T + 31) >> 5;
        const int want = (s_scan[0] / (nw > 0 ? nw : 1)) & 3;
        for (int w = 0; w < nw; w++)
            if ((s_scan[w] & 3) == want) {
                s_sweep_warp = w;
                break;
            }
    }
    __syncthreads();
    const int sweep_warp = s_sweep_warp;
    // The launches of the larger bands let the next one start right away (programmatic dependent launch: they have no
    // data dependency, they drain disjoint parts of the queue and only meet in integer atomics); every later launch
    // waits for the grid before it when it runs out of work, so that whatever follows in the stream sees all finished.
    if (p.role == 1 || p.role == 2) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int32_t *const q = p.job_meta + JOB_Q;
    const int n_jobs = (p.role == 1) ? q[1] : (p.role == 2) ? q[2] : q[0];
    int *const cursor = p.job_meta + JOB_Q + (p.role == 2 ? 4 : p.role == 3 ? 5 : 3);
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
            if (p.role >= 2) asm volatile("griddepcontrol.wait;" ::: "memory");
            break;
        }
        if (tid == 0) next_job = atomicAdd(cursor, 1);
        const int c0 = jt.x, n = jt.y, g = jt.z, cg = jt.w & 0x7fffffff;
        if (n <= 0) continue;  // empty, or the graph was skipped (arc buffer overflow)
        if (n <= p.rcap && jt.w >= 0 && staging_bytes(n, 64) <= p.sbytes)
            solve_component<true, (BT >= 512 ? 3 : 2)>(p, cg, c0, n, g, s_scan, s_red, s_ok, s_work, sweep_warp);
        else
            solve_component<false, 1>(p, cg, c0, n, g, s_scan, s_red, s_ok, s_work, sweep_warp);
    }
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
        const int pq = p.ws_par[q];
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

// ---- per-device facts and launch parameters are looked up once, not on every call
struct DevInfo {
    int dev, n_sm, smem_sm, smem_block;
    size_t smem_raised[3];  // dynamic shared memory the three instances of k_ssp_solve are allowed to use so far
    int occ_threads[3];
    size_t occ_smem[3];
    int occ_value[3];
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
    g_dev_known[dev] = true;
    return d;
}
static const void *ssp_instance(int which) {
    return which == 2 ? (const void *)k_ssp_solve<512, 1> : which == 1 ? (const void *)k_ssp_solve<256, 3>
                                                                        : (const void *)k_ssp_solve<128, 8>;
}
static int32_t raise_smem(const DevInfo *di, int which, size_t bytes) {
    std::lock_guard<std::mutex> lock(g_dev_mutex);
    DevInfo *d = &g_dev[di->dev];
    if (bytes <= d->smem_raised[which]) return MOT3D_OK;
    M3_CUDA(cudaFuncSetAttribute(ssp_instance(which), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    d->smem_raised[which] = bytes;
    return MOT3D_OK;
}
static int occupancy_of(const DevInfo *di, int which, int threads, size_t smem) {
    std::lock_guard<std::mutex> lock(g_dev_mutex);
    DevInfo *d = &g_dev[di->dev];
    if (d->occ_threads[which] == threads && d->occ_smem[which] == smem && d->occ_value[which] > 0) return d->occ_value[which];
    int v = 1;
    cudaError_t e = which == 2 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, k_ssp_solve<512, 1>, threads, smem)
                    : which == 1 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, k_ssp_solve<256, 3>, threads, smem)
                                 : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, k_ssp_solve<128, 8>, threads, smem);
    if (e != cudaSuccess) {
        cudaGetLastError();
        v = 1;
    }
    if (v < 1) v = 1;
    d->occ_threads[which] = threads;
    d->occ_smem[which] = smem;
    d->occ_value[which] = v;
    return v;
}
// Launch-shape experiments (bench sweeps) are compiled in with -DMOT3D_TUNING only; a production build has no knobs.
struct Tuning {
    int resident = 0, smem = 0, blocks_per_sm = 0, giant_rcap = 0;
    bool no_fast = false, no_giants = false;
};
static Tuning tuning() {
    Tuning t;
#ifdef MOT3D_TUNING
    static const Tuning cached = [] {
        Tuning c;
        if (const char *e = getenv("MOT3D_SOLVE_RESIDENT")) c.resident = atoi(e);
        if (const char *e = getenv("MOT3D_SOLVE_SMEM")) c.smem = atoi(e);
        if (const char *e = getenv("MOT3D_SOLVE_BLOCKS_PER_SM")) c.blocks_per_sm = atoi(e);
        if (const char *e = getenv("MOT3D_GIANT_RCAP")) c.giant_rcap = atoi(e);
        c.no_giants = getenv("MOT3D_SOLVE_NO_GIANTS") != nullptr;
        return c;
    }();
    t = cached;
#endif
    // test hook (tests/test_gpu_parity.py::test_single_track_pass_equals_full_solver): every component through the
    // full solver. Read on every call because the test flips it inside one process.
    t.no_fast = getenv("MOT3D_SOLVE_NO_FAST") != nullptr;
    return t;
}

}  // namespace m3

using namespace m3;

extern "C" {

size_t mot3d_solve_workspace_bytes(int64_t n_total, int32_t n_graphs, int64_t arc_capacity) {
    if (n_total < 0 || n_graphs < 0 || arc_capacity < 0) return 0;
    size_t n = (size_t)(n_total > 0 ? n_total : 1), g = (size_t)(n_graphs > 0 ? n_graphs : 1);
    size_t e = (size_t)(arc_capacity > 0 ? arc_capacity : 1);
    return a256(n * 8) * 3 + a256(n * 4) * 12 + a256(n * 16) + a256((n + g) * 4) * 2 + a256(g * 4) * 2 +
           a256(n * sizeof(RecT<false>)) + a256((e + n + 8) * sizeof(StArc)) +
           a256(mot3d_components_workspace_bytes(n_total, n_graphs)) + 512 + a256(JOB_META_INTS * sizeof(int32_t));
}

int32_t mot3d_solve_batch(int32_t n_graphs, int64_t n_total, const int32_t *graph_ptr, int32_t max_graph_nodes,
                          const int32_t *tkey, int64_t arc_capacity, const int32_t *in_ptr, const int32_t *in_src,
                          const int64_t *in_cost, const int64_t *entry_cost, const int64_t *node_cost,
                          const int64_t *exit_cost, int32_t *next_out, int32_t *prev_out, int32_t *n_paths_out,
                          int64_t *total_cost_out, int64_t *path_cost_out, int64_t *stats_out, void *ws, size_t ws_bytes,
                          void *stream) {
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
    p.rcap = 0;
    p.sbytes = 0;
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
    p.ws_dpre = (int64_t *)take(n_cap * 8);
    p.ws_dpost = (int64_t *)take(n_cap * 8);
    p.first_cost = (int64_t *)take(n_cap * 8);
    p.ws_par = (int32_t *)take(n_cap * 4);
    p.ws_anc_pre = (int32_t *)take(n_cap * 4);
    p.ws_anc_post = (int32_t *)take(n_cap * 4);
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
    p.ws_rec = (RecT<false> *)take(n_cap * sizeof(RecT<false>));
    p.ws_arc = (StArc *)take((e_cap + n_cap + 8) * sizeof(StArc));
    const size_t cws_bytes = mot3d_components_workspace_bytes(n_total, n_graphs);
    void *cws = take(cws_bytes);
    int32_t *n_cgraphs = (int32_t *)take(16);
    p.job_meta = (int32_t *)take(JOB_META_INTS * sizeof(int32_t));
    p.comp_class = (int32_t *)take(n_cap * 4);
    p.comp_perm = comp_perm;
    p.comp_inv = comp_inv;
    p.cgraph_ptr = cgraph_ptr;
    p.cgraph_win = cgraph_win;
    p.win_cbase = win_cbase;
    p.n_cgraphs = n_cgraphs;

    cudaStream_t st = (cudaStream_t)stream;
    M3_CUDA(cudaMemsetAsync(p.job_meta, 0, JOB_META_INTS * sizeof(int32_t), st));
    // 1. connected components of every graph, detections regrouped by component
    int32_t rc = components_run(n_graphs, n_total, graph_ptr, max_graph_nodes, in_ptr, in_src, arc_capacity, comp_perm,
                                comp_inv, cgraph_ptr, cgraph_win, win_cbase, n_cgraphs, cws, cws_bytes, stream);
    if (rc) return rc;
    // 2. first shortest-path tree per graph: 2 lanes per node, 256 threads: a layer of ~100 detections is one pass of
    //    the block
    k_first_tree<2><<<n_graphs, 256, 0, st>>>(p);
    M3_LAUNCH_CHECK("k_first_tree");
    // 3. successive shortest paths per component. Shared memory per block = the staging area of one component; blocks
    //    pull components from the work queue, which is cut into three bands by component size, every band drained by a
    //    launch whose staging area fits its components:
    //      band 0  up to 256 records (448 for batches of small graphs), 128 threads, 8 (4) blocks per SM
    //      band 1  up to 512 records, 256 threads, 3 blocks per SM ("giants": C5 windows, 257-384 detections)
    //      band 2  up to 1 536 records, 512 threads, one block per SM (long windows with crossing tracks)
    //    Larger components and components with costs beyond 32 bits run on the global staging area.
    const DevInfo *di = dev_info();
    if (!di) return cuda_fail(cudaGetLastError(), "device attributes");
    const int n_sm = di->n_sm, smem_sm = di->smem_sm;
    const Tuning tn = tuning();
    // Staging areas of the three launches: an eighth of an SM's shared memory (1 KB per block is reserved by the
    // system, ~1 KB is static), a third, all of it. Records are limited by the relabelling rounds: 2 per thread of
    // the 128- and 256-thread blocks, 3 per thread of the 512-thread blocks.
    const int32_t cap_rec[3] = {256, 512, 1536};
    int32_t cap_bytes[3];
    cap_bytes[0] = (smem_sm / 8 - 2048) & ~15;
    cap_bytes[1] = (smem_sm / 3 - 2048) & ~15;
    cap_bytes[2] = (di->smem_block - 2048) & ~15;
    for (int b = 0; b < 3; b++)
        if (cap_bytes[b] > di->smem_block) cap_bytes[b] = di->smem_block & ~15;
    // a batch of small graphs: a graph of up to 64 detections fits band 0 whatever its arcs (64 x 63 / 2 of them)
    const bool one_band = max_graph_nodes <= 64;
    p.band_rec[0] = one_band ? 0x7fffffff : cap_rec[0];
    p.band_bytes[0] = one_band ? 0x7fffffff : cap_bytes[0];
    p.band_rec[1] = cap_rec[1];
    p.band_bytes[1] = cap_bytes[1];
    p.rcap = cap_rec[0];
    p.sbytes = cap_bytes[0];
    p.fast = tn.no_fast ? 0 : 1;
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
    bool chained = false;  // a launch before this one lets it start early (programmatic dependent launch)
    auto launch = [&](int which, const SolveParams &pp, long long blocks, int threads, size_t bytes) -> int32_t {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3((unsigned)blocks);
        cfg.blockDim = dim3((unsigned)threads);
        cfg.dynamicSmemBytes = bytes;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = chained ? 1 : 0;
        chained = true;
        if (which == 2) M3_CUDA(cudaLaunchKernelEx(&cfg, k_ssp_solve<512, 1>, pp));
        if (which == 1) M3_CUDA(cudaLaunchKernelEx(&cfg, k_ssp_solve<256, 3>, pp));
        if (which == 0) M3_CUDA(cudaLaunchKernelEx(&cfg, k_ssp_solve<128, 8>, pp));
        return MOT3D_OK;
    };
    const int bt[3] = {128, 256, 512};
    for (int b = one_band ? 0 : 2; b >= 0; b--) {
        SolveParams pb = p;
        pb.role = one_band ? 0 : 3 - b;  // 1 = band 2 ... 3 = band 0
        pb.rcap = cap_rec[b];
        pb.sbytes = cap_bytes[b];
        if (cap_bytes[b] > 48 * 1024 && (rc = raise_smem(di, b, (size_t)cap_bytes[b]))) return rc;
        int per_sm = occupancy_of(di, b, bt[b], (size_t)cap_bytes[b]);
        if (b == 0 && tn.blocks_per_sm >= 1 && tn.blocks_per_sm < per_sm) per_sm = tn.blocks_per_sm;
        long long grid = (long long)n_sm * per_sm;
        // never more blocks than components the band can hold: band b > 0 only takes components that did not fit band
        // b - 1, i.e. of more than 64 detections
        const long long most = b == 0 ? n_total : n_total / 65 + 1;
        if (grid > most) grid = most > 0 ? most : 1;
        if ((rc = launch(b, pb, grid, bt[b], (size_t)cap_bytes[b]))) return rc;
    }
    M3_LAUNCH_CHECK("k_ssp_solve");
    // 4. the reference's "first path is always kept" rule, per graph
    k_first_path_rule<<<(n_graphs + 127) / 128, 128, 0, st>>>(p);
    M3_LAUNCH_CHECK("k_first_path_rule");
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
