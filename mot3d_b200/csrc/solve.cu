// Min-cost-flow solve: successive shortest paths on the residual network, one thread block per graph.
//
// Replaces wrapper.solve (reference mot3d/solvers/wrappers/muSSP/wrapper.cpp:171-317) and the muSSP core it drives
// (zip!/muSSP-master/muSSP/Graph.cpp).  Same network, same path-update rule, same acceptance rule, exact int64:
//   network        S -> pre_i (entry), pre_i -> post_i (node cost), post_i -> T (exit), post_i -> pre_j (transition);
//                  unit capacities                                   (mot3d/mot3d.py:75-141)
//   first tree     one relaxation pass over the frame layers         (Graph::shortest_path_dag, Graph.cpp:139-187)
//   next paths     muSSP's observation (Graph.cpp:341-343, 394-503): after a path through S -> pre_a is flipped,
//                  only the branch of the shortest-path tree that hung from that arc loses its labels; every
//                  other node keeps its exact distance.  The branch is collected (in frame-layer order) into
//                  shared memory together with its in-arcs and re-labelled by an exact Bellman-Ford fix-point in
//                  pull mode: ascending layer order settles the forward arcs, descending order settles the
//                  reversed arcs (they only run backwards in time along tracks); the two alternate until nothing
//                  moves.  No tolerances, no heap.
//   path update    reverse the arcs of the path                      (Graph::flip_path, Graph.cpp:266-335); with unit
//                  capacities the residual state is just next[]/prev[] per detection
//   acceptance     path 1 always; path k >= 2 while its cost < 0     (wrapper.cpp:199-267 on integers)
// Flow read-out (make_output2, wrapper.cpp:139-169) is unnecessary: next[]/prev[] are the flow.
#include "common.cuh"

namespace m3 {

constexpr int LPN = 4;  // lanes cooperating on one node's in-arcs

// state markers (local indices are >= 0)
constexpr int NONE = -1;  // detection unused
constexpr int TERM = -2;  // next: track ends here (post->T saturated); prev: track starts here (S->pre saturated)
constexpr int PAR_S = -1;     // pre reached from S
constexpr int PAR_POST = -2;  // pre reached from its own post (reversed node arc)
constexpr int PAR_NONE = -3;
constexpr int DL_PRE = 1 << 29;  // work-record flags: the detection's pre / post node is to be re-labelled
constexpr int DL_POST = 1 << 30;
constexpr int DL_MASK = DL_PRE - 1;

// One detection of the branch being re-labelled, with everything the sweeps need so that they never leave
// shared memory: node id + flags, frame layer key, where its in-arcs sit, its S->pre and pre->post costs.
struct __align__(16) Rec {
    int32_t jf;   // detection | DL_PRE | DL_POST
    int32_t tk;   // time key (frame layer)
    int32_t a0;   // first in-arc in the staged arc arrays, or -(1 + global arc index) when they did not fit
    int32_t deg;  // number of in-arcs (0 unless DL_PRE)
    int64_t en;
    int64_t cn;
};

struct SolveParams {
    int32_t n_graphs;
    int32_t rec_cap;  // work records / staged arcs that fit in shared memory
    int32_t arc_cap;
    int32_t pad;
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
    int64_t *path_cost;
    int64_t *stats;        // optional [n_graphs][4]: ascending sweeps, descending sweeps, node pulls, arc pulls
    int64_t arc_capacity;  // arcs actually present in in_src/in_cost (graphs reaching beyond it are skipped)
    // global workspace
    int64_t *ws_dpre;
    int64_t *ws_dpost;
    int64_t *ws_mcost;
    int32_t *ws_idx;  // 5 index arrays of n_total each (global-state mode)
    int32_t *ws_lvl;  // [n_total + n_graphs]
    Rec *ws_rec;      // [n_total] work records when the branch does not fit in shared memory
    int32_t *ws_fix;  // [n_total]
    int64_t n_total;
};

// Per-graph solver state. Distances are exact shortest distances from S in the current residual network for EVERY
// node (the invariant the incremental update relies on); anc_* = the detection whose S->pre arc starts the node's
// path in the shortest-path tree (its "branch"; muSSP's ancestor_node_id, Graph.cpp:160-161).
template <typename IdxT>
struct State {
    int64_t *dpre, *dpost, *mcost;
    IdxT *par, *nxt, *prv, *anc_pre, *anc_post;
    int32_t *asrc;  // staged in-arcs of the branch: source (local index) and cost
    int64_t *acost;
    const int32_t *in_src;  // global CSR (fallback when the branch's arcs did not fit)
    const int64_t *in_cost;
    int g0;
};

// pull for post_i: its only residual in-arc is pre_i -> post_i (detection unused) or pre_next -> post_i (reversed
// transition arc, cost -c). One lane.
template <typename IdxT>
__device__ __forceinline__ bool pull_post(const State<IdxT> &s, int i, int64_t cn_i) {
    int64_t c = INF;
    int a = -1;
    const int pi = s.prv[i];
    if (pi == NONE) {
        const int64_t d = s.dpre[i];
        if (d < INF_CUT) {
            c = d + cn_i;
            a = s.anc_pre[i];
        }
    } else {
        const int j = s.nxt[i];
        if (j >= 0) {
            const int64_t d = s.dpre[j];
            if (d < INF_CUT) {
                c = d - s.mcost[j];
                a = s.anc_pre[j];
            }
        }
    }
    if (c < s.dpost[i]) {
        s.dpost[i] = c;
        s.anc_post[i] = (IdxT)a;
        return true;
    }
    return false;
}

// finish the pull for pre_j given the best forward in-arc found by the lane group; adds S -> pre_j and the reversed
// node arc post_j -> pre_j. One lane.
template <typename IdxT>
__device__ __forceinline__ bool finish_pre(const State<IdxT> &s, int j, KeyIdx best, bool with_forward, int64_t en_j,
                                           int64_t cn_j) {
    const int pj = s.prv[j];
    int anc = -1;
    if (with_forward) {
        if (best.idx >= 0) anc = s.anc_post[best.idx];
        if (pj != TERM && en_j < best.key) {
            best.key = en_j;
            best.idx = PAR_S;
            anc = j;
        }
    }
    if (pj != NONE) {
        const int64_t d = s.dpost[j];
        if (d < INF_CUT) {
            const int64_t c = d - cn_j;
            if (c < best.key) {
                best.key = c;
                best.idx = PAR_POST;
                anc = s.anc_post[j];
            }
        }
    }
    if (best.key < s.dpre[j]) {
        s.dpre[j] = best.key;
        s.par[j] = (IdxT)best.idx;
        s.anc_pre[j] = (IdxT)anc;
        return true;
    }
    return false;
}

// best forward in-arc of pre_j seen by this lane (the lanes of a group stride over the row).
// a0 >= 0: staged arcs in shared memory; a0 < 0: -(1 + index) into the global CSR.
template <typename IdxT>
__device__ __forceinline__ KeyIdx scan_in_arcs(const State<IdxT> &s, int j, int a0, int deg, int sub) {
    KeyIdx best;
    best.key = INF;
    best.idx = PAR_NONE;
    const int pj = s.prv[j];
    if (a0 >= 0) {
        for (int k = sub; k < deg; k += LPN) {
            const int i = s.asrc[a0 + k];
            const int64_t di = s.dpost[i];
            if (i != pj && di < INF_CUT) {
                const int64_t cand = di + s.acost[a0 + k];
                if (cand < best.key) {  // sources ascend, so the first minimum has the lowest index
                    best.key = cand;
                    best.idx = i;
                }
            }
        }
    } else {
        const int e0 = -(a0 + 1);
        for (int k = sub; k < deg; k += LPN) {
            const int i = s.in_src[e0 + k] - s.g0;
            const int64_t di = s.dpost[i];
            if (i != pj && di < INF_CUT) {
                const int64_t cand = di + s.in_cost[e0 + k];
                if (cand < best.key) {
                    best.key = cand;
                    best.idx = i;
                }
            }
        }
    }
    return best;
}

__device__ __forceinline__ KeyIdx group_min(KeyIdx best) {
#pragma unroll
    for (int d = 1; d < LPN; d <<= 1) {
        KeyIdx o = shfl_xor_ki(best, d);
        if (o.key < best.key || (o.key == best.key && o.idx >= 0 && (best.idx < 0 || o.idx < best.idx))) best = o;
    }
    return best;
}

template <typename IdxT, bool SMEM_STATE>
__global__ void k_ssp_solve(const SolveParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_scan[33];
    __shared__ KeyIdx s_red[33];
    __shared__ int s_nfix, s_root;

    const int g = blockIdx.x;
    const int g0 = p.graph_ptr[g];
    const int n = p.graph_ptr[g + 1] - g0;
    const int T = blockDim.x;
    const int tid = threadIdx.x;
    if (n <= 0 || (int64_t)p.in_ptr[g0 + n] > p.arc_capacity) {
        // empty graph, or the arc buffer overflowed during the build (reported by the fill kernel's flag)
        if (tid == 0) {
            p.n_paths[g] = n <= 0 ? 0 : -1;
            p.total[g] = 0;
        }
        for (int i = tid; i < n; i += T) {
            p.next_out[g0 + i] = NONE;
            p.prev_out[g0 + i] = NONE;
        }
        return;
    }
    State<IdxT> s;
    unsigned char *sm = smem_raw;
    if (SMEM_STATE) {
        s.dpre = (int64_t *)sm;
        s.dpost = s.dpre + n;
        s.mcost = s.dpost + n;
        s.par = (IdxT *)(s.mcost + n);
        s.nxt = s.par + n;
        s.prv = s.nxt + n;
        s.anc_pre = s.prv + n;
        s.anc_post = s.anc_pre + n;
        sm = (unsigned char *)(((uintptr_t)(s.anc_post + n) + 15) & ~(uintptr_t)15);
    } else {
        s.dpre = p.ws_dpre + g0;
        s.dpost = p.ws_dpost + g0;
        s.mcost = p.ws_mcost + g0;
        s.par = (IdxT *)(p.ws_idx + g0);
        s.nxt = (IdxT *)(p.ws_idx + p.n_total + g0);
        s.prv = (IdxT *)(p.ws_idx + 2 * p.n_total + g0);
        s.anc_pre = (IdxT *)(p.ws_idx + 3 * p.n_total + g0);
        s.anc_post = (IdxT *)(p.ws_idx + 4 * p.n_total + g0);
    }
    // staging area for the branch being re-labelled
    Rec *rec_sm = (Rec *)sm;
    s.acost = (int64_t *)(rec_sm + p.rec_cap);
    s.asrc = (int32_t *)(s.acost + p.arc_cap);
    s.in_src = p.in_src;
    s.in_cost = p.in_cost;
    s.g0 = g0;
    Rec *rec_gl = p.ws_rec + g0;
    int32_t *lvl = p.ws_lvl + g0 + g;
    int32_t *fix = p.ws_fix + g0;
    const int32_t *tkey = p.tkey + g0;
    const int32_t *in_ptr = p.in_ptr + g0;
    const int64_t *en = p.en + g0, *cn = p.cn + g0, *ex = p.ex + g0;

    // ---- prologue: empty flow, nothing labelled, frame layers
    for (int i = tid; i < n; i += T) {
        s.nxt[i] = (IdxT)NONE;
        s.prv[i] = (IdxT)NONE;
        s.dpre[i] = INF;
        s.dpost[i] = INF;
        s.par[i] = (IdxT)PAR_NONE;
        s.anc_pre[i] = (IdxT)-1;
        s.anc_post[i] = (IdxT)-1;
    }
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

    int K = 0;
    int64_t total = 0;
    int64_t n_asc = 0, n_desc = 0, n_node = 0, n_arc = 0;
    const int sub = tid % LPN;
    const int grp = tid / LPN;
    const int ngrp = T / LPN;

    // ---- first tree: no flow yet, the network is a DAG: one sweep over the frame layers with the whole block
    //      (Graph::shortest_path_dag, Graph.cpp:139-187)
    for (int l = 0; l < L; l++) {
        const int a = lvl[l], b = lvl[l + 1];
        for (int jb = a; jb < b; jb += ngrp) {
            const int j = jb + grp;
            const bool valid = j < b;
            KeyIdx best;
            best.key = INF;
            best.idx = PAR_NONE;
            if (valid) {
                const int e0 = in_ptr[j];
                best = scan_in_arcs(s, j, -(e0 + 1), in_ptr[j + 1] - e0, sub);
            }
            best = group_min(best);
            if (valid && sub == 0) {
                const int64_t cn_j = cn[j];
                finish_pre(s, j, best, true, en[j], cn_j);
                pull_post(s, j, cn_j);
            }
        }
        __syncthreads();
    }
    n_asc = 1;
    n_node = 2 * (int64_t)n;
    const int64_t first_arcs = in_ptr[n] - in_ptr[0];

    for (int iter = 0; iter <= n; iter++) {  // at most one path per detection
        // ---- cheapest way into T (Graph::update_sink_info's job, Graph.cpp:508-541; here a block-wide arg-min)
        KeyIdx best;
        best.key = INF;
        best.idx = -1;
        for (int i = tid; i < n; i += T) {
            const int64_t d = s.dpost[i], x = ex[i];
            if (s.nxt[i] != TERM && d < INF_CUT && x < INF_CUT) {
                KeyIdx c;
                c.key = d + x;
                c.idx = i;
                best = min_ki(best, c);
            }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            KeyIdx o = shfl_xor_ki(best, d);
            if (o.idx >= 0 && (best.idx < 0 || o.key < best.key || (o.key == best.key && o.idx < best.idx))) best = o;
        }
        if (lane_id() == 0) s_red[warp_id()] = best;
        __syncthreads();
        if (warp_id() == 0) {
            const int nw = (T + 31) >> 5;
            KeyIdx b2;
            b2.key = INF;
            b2.idx = -1;
            if (lane_id() < nw) b2 = s_red[lane_id()];
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                KeyIdx o = shfl_xor_ki(b2, d);
                if (o.idx >= 0 && (b2.idx < 0 || o.key < b2.key || (o.key == b2.key && o.idx < b2.idx))) b2 = o;
            }
            if (lane_id() == 0) s_red[32] = b2;
        }
        __syncthreads();
        const KeyIdx sink = s_red[32];
        if (sink.idx < 0) break;             // T unreachable
        if (K >= 1 && sink.key >= 0) break;  // wrapper.cpp:263-267: stop when the next path does not pay
        // ---- augment: walk the path back from T and rewrite next/prev (Graph::flip_path, Graph.cpp:266-335)
        if (tid == 0) {
            int nfix = 0;
            int nxt_new = TERM;
            int cur = sink.idx;
            int root = -1;
            for (int step = 0; step <= n; step++) {
                int m;
                if (s.prv[cur] == NONE) {  // post_cur came from pre_cur: detection cur joins a track
                    s.nxt[cur] = (IdxT)nxt_new;
                    m = cur;
                } else {  // post_cur came from pre_m, m = its old successor: the arc cur->m is cancelled
                    m = s.nxt[cur];
                    s.nxt[cur] = (IdxT)nxt_new;
                    while (m >= 0 && s.par[m] == PAR_POST) {  // pre_m came from post_m: m leaves its track
                        const int m2 = s.nxt[m];
                        s.nxt[m] = (IdxT)NONE;
                        s.prv[m] = (IdxT)NONE;
                        m = m2;
                    }
                }
                if (m < 0) break;  // cannot happen on a consistent residual state
                const int q = s.par[m];
                if (q < 0) {  // PAR_S (anything else cannot happen)
                    s.prv[m] = (IdxT)TERM;
                    root = m;
                    break;
                }
                s.prv[m] = (IdxT)q;
                fix[nfix++] = m;
                nxt_new = m;
                cur = q;
            }
            s_nfix = nfix;
            s_root = root;
            if (p.path_cost) p.path_cost[g0 + K] = sink.key;
        }
        __syncthreads();
        // cost of every newly matched arc prev[m] -> m (used by the reversed arcs)
        const int nfix = s_nfix;
        for (int t = tid; t < nfix; t += T) {
            const int m = fix[t];
            const int src = s.prv[m] + g0;
            const int e1 = in_ptr[m + 1];
            for (int e = in_ptr[m]; e < e1; e++)
                if (p.in_src[e] == src) {
                    s.mcost[m] = p.in_cost[e];
                    break;
                }
        }
        K++;
        total += sink.key;
        const int root = s_root;
        if (root < 0) break;
        // ---- only the branch of the tree that hung from S -> pre_root lost its labels
        //      (Graph::find_node_set4update, Graph.cpp:341-343). Pass 1: count it.
        int nd = 0;
        for (int base = 0; base < n; base += T) {
            const int i = base + tid;
            const int f = (i < n && (s.anc_pre[i] == root || s.anc_post[i] == root)) ? 1 : 0;
            nd += __syncthreads_count(f);
        }
        Rec *rec = (nd <= p.rec_cap) ? rec_sm : rec_gl;
        // Pass 2: work records in layer order (ascending detection index), labels reset, arc offsets by a scan of
        // the in-degrees; arcs that do not fit in shared memory are read from the global CSR by the sweeps.
        int pos = 0, aoff = 0;
        for (int base = 0; base < n; base += T) {
            const int i = base + tid;
            int f = 0, deg = 0, e0 = 0;
            if (i < n) {
                if (s.anc_pre[i] == root) f |= DL_PRE;
                if (s.anc_post[i] == root) f |= DL_POST;
                if (f & DL_PRE) {
                    e0 = in_ptr[i];
                    deg = in_ptr[i + 1] - e0;
                }
            }
            int tot, atot;
            const int slot = pos + block_excl_scan(f ? 1 : 0, s_scan, &tot);
            const int a0 = aoff + block_excl_scan(deg, s_scan, &atot);
            if (f) {
                Rec r;
                r.jf = i | f;
                r.tk = tkey[i];
                r.deg = deg;
                r.a0 = (a0 + deg <= p.arc_cap) ? a0 : -(e0 + 1);
                r.en = en[i];
                r.cn = cn[i];
                rec[slot] = r;
                if (r.a0 >= 0)
                    for (int k = 0; k < deg; k++) {
                        s.asrc[a0 + k] = p.in_src[e0 + k] - g0;
                        s.acost[a0 + k] = p.in_cost[e0 + k];
                    }
                if (f & DL_PRE) {
                    s.dpre[i] = INF;
                    s.par[i] = (IdxT)PAR_NONE;
                    s.anc_pre[i] = (IdxT)-1;
                }
                if (f & DL_POST) {
                    s.dpost[i] = INF;
                    s.anc_post[i] = (IdxT)-1;
                }
            }
            pos += tot;
            aoff += atot;
        }
        __syncthreads();
        // ---- re-label the branch: exact Bellman-Ford fix-point over its nodes, warp 0 only (few nodes per layer).
        if (warp_id() == 0) {
            const int lane = lane_id();
            constexpr int NPC = 32 / LPN;  // nodes per ascending chunk
            for (int round = 0; round < 2 * n + 4; round++) {
                bool ch = false;
                n_asc++;
                for (int q = 0; q < nd;) {
                    // chunk = up to NPC consecutive records of one layer, LPN lanes each
                    Rec r;
                    r.jf = 0;
                    r.tk = 0;
                    r.a0 = 0;
                    r.deg = 0;
                    r.en = 0;
                    r.cn = 0;
                    const int my = lane / LPN;
                    const bool in_range = q + my < nd;
                    if (in_range) r = rec[q + my];
                    const int tk0 = __shfl_sync(0xffffffffu, r.tk, 0);
                    const unsigned same = __ballot_sync(0xffffffffu, in_range && r.tk == tk0);
                    // lanes of group k occupy bits [k*LPN, (k+1)*LPN); count the leading groups of the same layer
                    int cnt = 0;
#pragma unroll
                    for (int k = 0; k < NPC; k++)
                        if (cnt == k && ((same >> (k * LPN)) & 1u)) cnt = k + 1;
                    const bool valid = my < cnt;
                    const int j = r.jf & DL_MASK;
                    KeyIdx best;
                    best.key = INF;
                    best.idx = PAR_NONE;
                    if (valid && (r.jf & DL_PRE)) best = scan_in_arcs(s, j, r.a0, r.deg, lane % LPN);
                    best = group_min(best);
                    if (valid && lane % LPN == 0) {
                        if (r.jf & DL_PRE) {
                            ch |= finish_pre(s, j, best, true, r.en, r.cn);
                            n_arc += r.deg;
                        }
                        if (r.jf & DL_POST) ch |= pull_post(s, j, r.cn);
                    }
                    n_node += cnt;
                    q += cnt;
                    __syncwarp();
                }
                if (!__any_sync(0xffffffffu, ch) && round > 0) break;
                bool ch2 = false;
                n_desc++;
                for (int q = nd; q > 0;) {
                    // chunk = up to 32 consecutive records of one layer, one lane each (O(1) work per node)
                    Rec r;
                    r.jf = 0;
                    r.tk = 0;
                    r.a0 = 0;
                    r.deg = 0;
                    r.en = 0;
                    r.cn = 0;
                    const bool in_range = q - 1 - lane >= 0;
                    if (in_range) r = rec[q - 1 - lane];
                    const int tk0 = __shfl_sync(0xffffffffu, r.tk, 0);
                    const unsigned same = __ballot_sync(0xffffffffu, in_range && r.tk == tk0);
                    const int cnt = (same == 0xffffffffu) ? 32 : __ffs(~same) - 1;
                    if (lane < cnt) {
                        const int j = r.jf & DL_MASK;
                        if (r.jf & DL_POST) ch2 |= pull_post(s, j, r.cn);
                        KeyIdx none;
                        none.key = INF;
                        none.idx = PAR_NONE;
                        if (r.jf & DL_PRE) ch2 |= finish_pre(s, j, none, false, r.en, r.cn);
                    }
                    n_node += cnt;
                    q -= cnt;
                    __syncwarp();
                }
                if (!__any_sync(0xffffffffu, ch2)) break;
            }
        }
        __syncthreads();
    }
    // ---- results
    if (p.stats && warp_id() == 0) {
        // arc pulls of the sweeps were counted by the group leaders of warp 0
        int64_t v = n_arc;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        if (tid == 0) {
            p.stats[4 * g + 0] = n_asc;
            p.stats[4 * g + 1] = n_desc;
            p.stats[4 * g + 2] = n_node;
            p.stats[4 * g + 3] = v + first_arcs;
        }
    }
    if (tid == 0) {
        p.n_paths[g] = K;
        p.total[g] = total;
    }
    for (int i = tid; i < n; i += T) {
        const int a = s.nxt[i], b = s.prv[i];
        p.next_out[g0 + i] = a >= 0 ? a + g0 : a;
        p.prev_out[g0 + i] = b >= 0 ? b + g0 : b;
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

constexpr size_t SMEM_BUDGET = 227 * 1024 - 1024;  // dynamic shared memory we ask for at most (static use ~700 B)

}  // namespace m3

using namespace m3;

extern "C" {

size_t mot3d_solve_workspace_bytes(int64_t n_total, int32_t n_graphs) {
    if (n_total < 0 || n_graphs < 0) return 0;
    size_t n = (size_t)(n_total > 0 ? n_total : 1), g = (size_t)(n_graphs > 0 ? n_graphs : 1);
    return a256(n * 8) * 3 + a256(n * 4 * 5) + a256((n + g) * 4) + a256(n * sizeof(Rec)) + a256(n * 4) + 256;
}

int32_t mot3d_solve_batch(int32_t n_graphs, int64_t n_total, const int32_t *graph_ptr, int32_t max_graph_nodes,
                          const int32_t *tkey, int64_t arc_capacity, const int32_t *in_ptr, const int32_t *in_src,
                          const int64_t *in_cost, const int64_t *entry_cost, const int64_t *node_cost,
                          const int64_t *exit_cost, int32_t *next_out, int32_t *prev_out, int32_t *n_paths_out,
                          int64_t *total_cost_out, int64_t *path_cost_out, int64_t *stats_out, void *ws,
                          size_t ws_bytes, void *stream) {
    M3_CHECK_ARG(n_graphs >= 0 && max_graph_nodes >= 0, "negative sizes");
    if (n_graphs == 0) return MOT3D_OK;
    M3_CHECK_ARG(graph_ptr && tkey && in_ptr && entry_cost && node_cost && exit_cost, "NULL input");
    M3_CHECK_ARG(next_out && prev_out && n_paths_out && total_cost_out && ws, "NULL output/workspace");
    M3_CHECK_ARG(n_total >= max_graph_nodes, "n_total < max_graph_nodes");
    M3_CHECK_ARG(max_graph_nodes < DL_PRE, "graph too large");
    if (ws_bytes < mot3d_solve_workspace_bytes(n_total, n_graphs)) {
        set_error("solve workspace too small: %zu < %zu", ws_bytes, mot3d_solve_workspace_bytes(n_total, n_graphs));
        return MOT3D_E_CAPACITY;
    }
    const size_t g = (size_t)n_graphs;
    const size_t n_cap = (size_t)(n_total > 0 ? n_total : 1);
    char *w = (char *)ws;
    SolveParams p;
    p.n_graphs = n_graphs;
    p.pad = 0;
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
    p.n_total = (int64_t)n_cap;
    p.ws_dpre = (int64_t *)w;
    w += a256(n_cap * 8);
    p.ws_dpost = (int64_t *)w;
    w += a256(n_cap * 8);
    p.ws_mcost = (int64_t *)w;
    w += a256(n_cap * 8);
    p.ws_idx = (int32_t *)w;
    w += a256(n_cap * 4 * 5);
    p.ws_lvl = (int32_t *)w;
    w += a256((n_cap + g) * 4);
    p.ws_rec = (Rec *)w;
    w += a256(n_cap * sizeof(Rec));
    p.ws_fix = (int32_t *)w;

    // Shared memory plan: per-node state (24 B distances/costs + 5 indices of 2 or 4 B) when the largest graph fits,
    // the rest for the branch staging area (32 B records, 12 B arcs).
    const size_t nmax = (size_t)max_graph_nodes;
    const bool idx16 = nmax < 32000;
    const size_t state_bytes = nmax * (24 + 5 * (idx16 ? 2 : 4)) + 16;
    const size_t min_stage = 64 * sizeof(Rec) + 512 * 12;
    const bool smem_state = state_bytes + min_stage <= SMEM_BUDGET;
    size_t stage = smem_state ? SMEM_BUDGET - state_bytes : (size_t)96 * 1024;
    // small graphs: do not hog the SM, several blocks should fit
    const size_t want_stage = nmax * sizeof(Rec) / 2 + nmax * 6 * 12 + min_stage;
    if (stage > want_stage) stage = want_stage;
    size_t rec_cap = stage / 3 / sizeof(Rec);
    if (rec_cap > nmax) rec_cap = nmax;
    if (rec_cap < 64) rec_cap = 64;
    size_t arc_cap = (stage - rec_cap * sizeof(Rec)) / 12;
    p.rec_cap = (int32_t)rec_cap;
    p.arc_cap = (int32_t)arc_cap;
    const size_t smem = (smem_state ? state_bytes : 16) + rec_cap * sizeof(Rec) + arc_cap * 12 + 16;
    const int threads = max_graph_nodes <= 256 ? 128 : (max_graph_nodes <= 2048 ? 256 : 512);
    cudaStream_t st = (cudaStream_t)stream;
#define M3_LAUNCH_SOLVE(IDX, SM)                                                                                   \
    do {                                                                                                           \
        if (smem > 48 * 1024)                                                                                      \
            M3_CUDA(cudaFuncSetAttribute(k_ssp_solve<IDX, SM>, cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                                         (int)smem));                                                              \
        k_ssp_solve<IDX, SM><<<n_graphs, threads, smem, st>>>(p);                                                  \
    } while (0)
    if (smem_state && idx16)
        M3_LAUNCH_SOLVE(int16_t, true);
    else if (smem_state)
        M3_LAUNCH_SOLVE(int32_t, true);
    else
        M3_LAUNCH_SOLVE(int32_t, false);
#undef M3_LAUNCH_SOLVE
    M3_LAUNCH_CHECK("k_ssp_solve");
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
