// Min-cost-flow solve: successive shortest paths on the residual network, one thread block per graph.
//
// Replaces wrapper.solve (reference mot3d/solvers/wrappers/muSSP/wrapper.cpp:171-317) and the muSSP core it drives
// (zip!/muSSP-master/muSSP/Graph.cpp).  Same network, same path-update rule, same acceptance rule, exact int64:
//   network        S -> pre_i (entry), pre_i -> post_i (node cost), post_i -> T (exit), post_i -> pre_j (transition);
//                  unit capacities                                   (mot3d/mot3d.py:75-141)
//   first tree     one relaxation pass over the frame layers         (Graph::shortest_path_dag, Graph.cpp:139-187)
//   next paths     shortest S->T path in the residual network        (Graph.cpp:394-503), found here by frame-layered
//                  pull relaxation: a forward sweep over the layers settles every forward arc, a walk back along each
//                  existing track settles the reversed arcs (they only exist along tracks); the two alternate until
//                  nothing changes.  Exact Bellman-Ford fix-point, no tolerances.
//   path update    reverse the arcs of the path                      (Graph::flip_path, Graph.cpp:266-335); with unit
//                  capacities the residual state is just next[]/prev[] per detection
//   acceptance     path 1 always; path k >= 2 while its cost < 0     (wrapper.cpp:199-267 on integers)
// Flow read-out (make_output2, wrapper.cpp:139-169) is unnecessary: next[]/prev[] are the flow.
#include "common.cuh"

namespace m3 {

constexpr int LPN = 4;  // lanes cooperating on one node's in-arcs

struct SolveParams {
    int32_t n_graphs;
    int32_t use_smem;
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
    int32_t *stats;  // optional [n_graphs][4]: forward sweeps, track-walk phases, layers, arcs
    // global workspace
    int64_t *ws_dpre;
    int64_t *ws_dpost;
    int64_t *ws_mcost;
    int32_t *ws_par;
    int32_t *ws_lvl;   // [n_total + n_graphs]
    int32_t *ws_ends;  // [n_total]
    int32_t *ws_fix;   // [n_total]
};

// state markers (local indices are >= 0)
constexpr int NONE = -1;  // detection unused
constexpr int TERM = -2;  // next: track ends here (post->T saturated); prev: track starts here (S->pre saturated)
constexpr int PAR_S = -1;     // pre reached from S
constexpr int PAR_POST = -2;  // pre reached from its own post (reversed node arc)
constexpr int PAR_NONE = -3;

__global__ void k_ssp_solve(const SolveParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_scan[33];
    __shared__ KeyIdx s_red[33];
    __shared__ int s_nfix;

    const int g = blockIdx.x;
    const int g0 = p.graph_ptr[g];
    const int n = p.graph_ptr[g + 1] - g0;
    const int T = blockDim.x;
    const int tid = threadIdx.x;
    if (n <= 0) {
        if (tid == 0) {
            p.n_paths[g] = 0;
            p.total[g] = 0;
        }
        return;
    }
    int64_t *dpre, *dpost, *mcost;
    int32_t *par, *nxt, *prv;
    if (p.use_smem) {
        dpre = (int64_t *)smem_raw;
        dpost = dpre + n;
        mcost = dpost + n;
        par = (int32_t *)(mcost + n);
        nxt = par + n;
        prv = nxt + n;
    } else {
        dpre = p.ws_dpre + g0;
        dpost = p.ws_dpost + g0;
        mcost = p.ws_mcost + g0;
        par = p.ws_par + g0;
        nxt = p.next_out + g0;  // local indices while solving, globalised at the end
        prv = p.prev_out + g0;
    }
    int32_t *lvl = p.ws_lvl + g0 + g;
    int32_t *ends = p.ws_ends + g0;
    int32_t *fix = p.ws_fix + g0;
    const int32_t *tkey = p.tkey + g0;
    const int32_t *in_ptr = p.in_ptr + g0;
    const int64_t *en = p.en + g0, *cn = p.cn + g0, *ex = p.ex + g0;

    // ---- prologue: empty flow, frame layers
    for (int i = tid; i < n; i += T) {
        nxt[i] = NONE;
        prv[i] = NONE;
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
    int n_fwd = 0, n_bwd = 0;
    const int sub = tid % LPN;
    const int grp = tid / LPN;
    const int ngrp = T / LPN;

    for (int iter = 0; iter <= n; iter++) {  // at most one path per detection
        for (int i = tid; i < n; i += T) {
            dpre[i] = INF;
            dpost[i] = INF;
            par[i] = PAR_NONE;
        }
        __syncthreads();
        for (int round = 0; round < 4 * n + 8; round++) {  // fix-point; the bound only guards against bad input
            int changed = 0;
            n_fwd++;
            // ---- forward sweep over the frame layers (all forward arcs, pull mode)
            for (int l = 0; l < L; l++) {
                const int a = lvl[l], b = lvl[l + 1];
                for (int jb = a; jb < b; jb += ngrp) {
                    const int j = jb + grp;
                    const bool valid = j < b;
                    KeyIdx best;
                    best.key = INF;
                    best.idx = PAR_NONE;
                    int pj = NONE;
                    if (valid) {
                        pj = prv[j];
                        const int e1 = in_ptr[j + 1];
                        for (int e = in_ptr[j] + sub; e < e1; e += LPN) {
                            const int i = p.in_src[e] - g0;
                            const int64_t di = dpost[i];
                            if (i != pj && di < INF_CUT) {
                                const int64_t cand = di + p.in_cost[e];
                                if (cand < best.key) {  // sources ascend, so the first minimum has the lowest index
                                    best.key = cand;
                                    best.idx = i;
                                }
                            }
                        }
                    }
#pragma unroll
                    for (int d = 1; d < LPN; d <<= 1) {
                        KeyIdx o = shfl_xor_ki(best, d);
                        if (o.key < best.key || (o.key == best.key && o.idx >= 0 && o.idx < best.idx)) best = o;
                    }
                    if (valid && sub == 0) {
                        int64_t cur = dpre[j];
                        if (pj != TERM) {
                            const int64_t e_ = en[j];
                            if (e_ < best.key) {
                                best.key = e_;
                                best.idx = PAR_S;
                            }
                        }
                        if (best.key < cur) {
                            dpre[j] = best.key;
                            par[j] = best.idx;
                            cur = best.key;
                            changed = 1;
                        }
                        if (pj == NONE && cur < INF_CUT) dpost[j] = cur + cn[j];
                    }
                }
                __syncthreads();
            }
            if (K == 0) break;  // no flow yet: the network is a DAG, one sweep is exact
            if (!__syncthreads_or(changed) && round > 0) break;
            // ---- reversed arcs only exist along tracks: walk each track from its end to its start
            int ch2 = 0;
            n_bwd++;
            for (int t = tid; t < K; t += T) {
                int j = ends[t];
                for (int step = 0; step < n; step++) {
                    const int64_t dj = dpost[j];
                    if (dj < INF_CUT) {
                        const int64_t c = dj - cn[j];  // post_j -> pre_j, cost -node
                        if (c < dpre[j]) {
                            dpre[j] = c;
                            par[j] = PAR_POST;
                            ch2 = 1;
                        }
                    }
                    const int i = prv[j];
                    if (i < 0) break;
                    const int64_t dp = dpre[j];
                    if (dp < INF_CUT) {
                        const int64_t c = dp - mcost[j];  // pre_j -> post_i, cost -c_ij
                        if (c < dpost[i]) {
                            dpost[i] = c;
                            ch2 = 1;
                        }
                    }
                    j = i;
                }
            }
            if (!__syncthreads_or(ch2)) break;
        }
        // ---- cheapest way into T
        KeyIdx best;
        best.key = INF;
        best.idx = -1;
        for (int i = tid; i < n; i += T) {
            const int64_t d = dpost[i], x = ex[i];
            if (nxt[i] != TERM && d < INF_CUT && x < INF_CUT) {
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
        if (sink.idx < 0) break;              // T unreachable
        if (K >= 1 && sink.key >= 0) break;   // wrapper.cpp:263-267: stop when the next path does not pay
        // ---- augment: walk the path back from T and rewrite next/prev (Graph::flip_path)
        if (tid == 0) {
            int nfix = 0;
            int nxt_new = TERM;
            int cur = sink.idx;
            for (int step = 0; step <= n; step++) {
                int m;
                if (prv[cur] == NONE) {  // post_cur came from pre_cur: detection cur joins a track
                    nxt[cur] = nxt_new;
                    m = cur;
                } else {  // post_cur came from pre_m, m = its old successor: the arc cur->m is cancelled
                    m = nxt[cur];
                    nxt[cur] = nxt_new;
                    while (m >= 0 && par[m] == PAR_POST) {  // pre_m came from post_m: detection m leaves its track
                        const int m2 = nxt[m];
                        nxt[m] = NONE;
                        prv[m] = NONE;
                        m = m2;
                    }
                }
                if (m < 0) break;  // cannot happen on a consistent residual state
                const int q = par[m];
                if (q < 0) {  // PAR_S (anything else cannot happen)
                    prv[m] = TERM;
                    break;
                }
                prv[m] = q;
                fix[nfix++] = m;
                nxt_new = m;
                cur = q;
            }
            ends[K] = sink.idx;
            s_nfix = nfix;
            if (p.path_cost) p.path_cost[g0 + K] = sink.key;
        }
        __syncthreads();
        // cost of every newly matched arc prev[m] -> m (needed by the walks along tracks)
        const int nfix = s_nfix;
        for (int t = tid; t < nfix; t += T) {
            const int m = fix[t];
            const int src = prv[m] + g0;
            const int e1 = in_ptr[m + 1];
            for (int e = in_ptr[m]; e < e1; e++)
                if (p.in_src[e] == src) {
                    mcost[m] = p.in_cost[e];
                    break;
                }
        }
        K++;
        total += sink.key;
        __syncthreads();
    }
    // ---- results
    if (tid == 0) {
        p.n_paths[g] = K;
        p.total[g] = total;
        if (p.stats) {
            p.stats[4 * g + 0] = n_fwd;
            p.stats[4 * g + 1] = n_bwd;
            p.stats[4 * g + 2] = L;
            p.stats[4 * g + 3] = in_ptr[n] - in_ptr[0];
        }
    }
    for (int i = tid; i < n; i += T) {
        const int a = nxt[i], b = prv[i];
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

}  // namespace m3

using namespace m3;

extern "C" {

size_t mot3d_solve_workspace_bytes(int64_t n_total, int32_t n_graphs) {
    if (n_total < 0 || n_graphs < 0) return 0;
    size_t n = (size_t)(n_total > 0 ? n_total : 1), g = (size_t)(n_graphs > 0 ? n_graphs : 1);
    return a256(n * 8) * 3 + a256(n * 4) + a256((n + g) * 4) + a256(n * 4) * 2 + 256;
}

int32_t mot3d_solve_batch(int32_t n_graphs, int64_t n_total, const int32_t *graph_ptr, int32_t max_graph_nodes,
                          const int32_t *tkey,
                          const int32_t *in_ptr, const int32_t *in_src, const int64_t *in_cost,
                          const int64_t *entry_cost, const int64_t *node_cost, const int64_t *exit_cost,
                          int32_t *next_out, int32_t *prev_out, int32_t *n_paths_out, int64_t *total_cost_out,
                          int64_t *path_cost_out, int32_t *stats_out, void *ws, size_t ws_bytes, void *stream) {
    M3_CHECK_ARG(n_graphs >= 0 && max_graph_nodes >= 0, "negative sizes");
    if (n_graphs == 0) return MOT3D_OK;
    M3_CHECK_ARG(graph_ptr && tkey && in_ptr && entry_cost && node_cost && exit_cost, "NULL input");
    M3_CHECK_ARG(next_out && prev_out && n_paths_out && total_cost_out && ws, "NULL output/workspace");
    M3_CHECK_ARG(n_total >= max_graph_nodes, "n_total < max_graph_nodes");
    if (ws_bytes < mot3d_solve_workspace_bytes(n_total, n_graphs)) {
        set_error("solve workspace too small: %zu < %zu", ws_bytes, mot3d_solve_workspace_bytes(n_total, n_graphs));
        return MOT3D_E_CAPACITY;
    }
    const size_t g = (size_t)n_graphs;
    const size_t n_cap = (size_t)(n_total > 0 ? n_total : 1);
    char *w = (char *)ws;
    SolveParams p;
    p.n_graphs = n_graphs;
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
    p.ws_dpre = (int64_t *)w;
    w += a256(n_cap * 8);
    p.ws_dpost = (int64_t *)w;
    w += a256(n_cap * 8);
    p.ws_mcost = (int64_t *)w;
    w += a256(n_cap * 8);
    p.ws_par = (int32_t *)w;
    w += a256(n_cap * 4);
    p.ws_lvl = (int32_t *)w;
    w += a256((n_cap + g) * 4);
    p.ws_ends = (int32_t *)w;
    w += a256(n_cap * 4);
    p.ws_fix = (int32_t *)w;

    const size_t smem_need = (size_t)max_graph_nodes * 36 + 16;
    int threads = max_graph_nodes <= 256 ? 128 : (max_graph_nodes <= 2048 ? 256 : 512);
    cudaStream_t st = (cudaStream_t)stream;
    size_t smem = 0;
    p.use_smem = 0;
    if (smem_need <= 200 * 1024) {
        p.use_smem = 1;
        smem = smem_need;
        if (smem > 48 * 1024)
            M3_CUDA(cudaFuncSetAttribute(k_ssp_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    k_ssp_solve<<<n_graphs, threads, smem, st>>>(p);
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
