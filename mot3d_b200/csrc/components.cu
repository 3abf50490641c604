// Connected components of every graph of a batch, and the regrouping of its detections by component.
//
// Why: a tracking window decomposes into groups of detections that never come within max_distance/max_jump of each
// other; the min-cost flow of the window is the union of the min-cost flows of the groups (the arcs through S and T
// are the only coupling and the acceptance rule "keep a path while its cost < 0", wrapper.cpp:263-267, applies to
// every path on its own). Solving the groups as independent small graphs turns ~100 dependent augmentations per
// window into ~60 independent chains of 1-3, and the per-augmentation scans shrink from the window to the group.
// The reference has no counterpart (muSSP handles one graph sequentially); results are identical.
//
// One block per graph: a lock-free union-find over the transition arcs in shared memory (roots hook under smaller
// roots, so the root of a component is its smallest detection), then a stable counting sort by component rank gives
// the detections grouped by component, ascending detection index inside a component (= frame order, which the solver's
// layer sweeps need). Graphs larger than CC_MAX_SORT detections are kept whole.
#include "common.cuh"

extern "C" size_t mot3d_components_workspace_bytes(int64_t n_total, int32_t n_graphs);

namespace m3 {

constexpr int CC_TPB = 512;
constexpr int CC_MAX_SORT = 8192;

struct CompParams {
    int32_t n_graphs;
    int32_t g_first;       // first graph of this launch (block b works on graph g_first + b)
    int32_t cap;           // shared-memory capacity in detections (<= CC_MAX_SORT); larger graphs are kept whole
    const int32_t *graph_ptr;
    const int32_t *in_ptr;
    const int32_t *in_src;
    int64_t arc_capacity;
    int32_t *comp_perm;    // [n_total] out: detection (global index) at every regrouped position
    int32_t *comp_inv;     // [n_total] out: regrouped position (global) of every detection
    int32_t *win_ncomp;    // [n_graphs] out: components per graph
    int32_t *win_cstart;   // [n_total] out: graph g's component c starts at regrouped position win_cstart[g0 + c]
    FtOut tree;            // k_components_tree: the first tree, detection order in, regrouped order out
};

// Union-find on shared memory, safe under concurrent hooks: parents only ever decrease, so any value read is an
// ancestor and the racing path-halving writes are benign.
__device__ __forceinline__ int cc_find(volatile int *par, int x) {
    int q = par[x];
    while (q != x) {
        const int gq = par[q];
        if (gq != q) par[x] = gq;
        x = q;
        q = gq;
    }
    return x;
}
__device__ __forceinline__ int cc_unite(int *par, int ra, int rb) {  // returns a representative of the merged set
    while (ra != rb) {
        if (ra < rb) {
            const int t = ra;
            ra = rb;
            rb = t;
        }
        const int old = atomicCAS(&par[ra], ra, rb);  // hook the larger root under the smaller one
        if (old == ra) return rb;
        ra = cc_find(par, old);  // ra was hooked by somebody else meanwhile
    }
    return rb;
}

__global__ void __launch_bounds__(CC_TPB) k_components(const CompParams p) {
    extern __shared__ int s_dyn[];
    __shared__ int s_scan[33];
    int *s_par = s_dyn;          // [cap] parent -> root -> component rank -> regrouped position
    int *s_cnt = s_dyn + p.cap;  // [cap] rank of a root -> detections per component -> placement cursor
    const int g = blockIdx.x + p.g_first;
    const int g0 = p.graph_ptr[g];
    const int n = p.graph_ptr[g + 1] - g0;
    const int T = blockDim.x, tid = threadIdx.x;
    if (n <= 0) {
        if (tid == 0) p.win_ncomp[g] = 0;
        return;
    }
    const int32_t *in_ptr = p.in_ptr + g0;
    const bool overflow = (int64_t)in_ptr[n] > p.arc_capacity;  // arc buffer overflow: reported by the fill kernel
    const bool split = n <= p.cap && !overflow;
    if (!split) {  // kept whole: one component, identity order
        for (int i = tid; i < n; i += T) {
            p.comp_perm[g0 + i] = g0 + i;
            p.comp_inv[g0 + i] = g0 + i;
        }
        if (tid == 0) {
            p.win_ncomp[g] = 1;
            p.win_cstart[g0] = g0;
        }
        return;
    }
    for (int i = tid; i < n; i += T) s_par[i] = i;
    __syncthreads();
    // ---- one pass over the arcs: unite every detection with the sources of its incoming transitions. A warp takes
    //      32 consecutive rows: their arcs are one contiguous range, read with coalesced, independent loads; the row of
    //      an arc is found by a binary search over the 32 row starts held by the lanes.
    {
        const int lane = lane_id(), nwarp = T >> 5;
        for (int j0 = warp_id() * 32; j0 < n; j0 += nwarp * 32) {
            const int jl = j0 + lane;
            const int start = in_ptr[jl < n ? jl : n];  // rows beyond the graph: empty
            const int a0 = __shfl_sync(0xffffffffu, start, 0);
            const int a1 = in_ptr[j0 + 32 < n ? j0 + 32 : n];
            // the sources of the next two steps are already on their way while this step works on shared memory
            int src0 = a0 + lane < a1 ? p.in_src[a0 + lane] : 0;
            int src1 = a0 + 32 + lane < a1 ? p.in_src[a0 + 32 + lane] : 0;
            for (int eb = a0; eb < a1; eb += 32) {
                const int e = eb + lane;
                const int src2 = e + 64 < a1 ? p.in_src[e + 64] : 0;
                int lo = 0;  // the last row whose arcs start at or before e
#pragma unroll
                for (int step = 16; step > 0; step >>= 1) {
                    const int s = __shfl_sync(0xffffffffu, start, (lo + step) & 31);
                    if (s <= e) lo += step;
                }
                if (e < a1) {
                    const int ri = cc_find(s_par, src0 - g0);
                    const int rj = cc_find(s_par, j0 + lo);
                    cc_unite(s_par, ri, rj);
                }
                src0 = src1;
                src1 = src2;
            }
        }
    }
    __syncthreads();
    // flatten: every detection points at its root. Read-only chase (no path halving here: a halving write landing
    // after the owner's final write would leave a non-root behind); only the owner writes an entry.
    for (int i = tid; i < n; i += T) {
        volatile int *par = s_par;
        int r = par[i], q = par[r];
        while (q != r) {
            r = q;
            q = par[r];
        }
        par[i] = r;
    }
    __syncthreads();
    // ---- rank the roots (ascending root index = ascending first detection of the component)
    int ncomp = 0;
    for (int base = 0; base < n; base += T) {
        const int i = base + tid;
        const int f = (i < n && s_par[i] == i) ? 1 : 0;
        int tot;
        const int ex = block_excl_scan(f, s_scan, &tot);
        if (f) s_cnt[i] = ncomp + ex;
        ncomp += tot;
    }
    __syncthreads();
    for (int i = tid; i < n; i += T) s_par[i] = s_cnt[s_par[i]];  // own entry only: nobody else reads s_par[i] here
    __syncthreads();
    for (int c = tid; c < ncomp; c += T) s_cnt[c] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += T) atomicAdd(&s_cnt[s_par[i]], 1);
    __syncthreads();
    int run = 0;
    for (int base = 0; base < ncomp; base += T) {  // exclusive scan of the component sizes -> component starts
        const int c = base + tid;
        const int v = c < ncomp ? s_cnt[c] : 0;
        int tot;
        const int ex = block_excl_scan(v, s_scan, &tot);
        if (c < ncomp) {
            s_cnt[c] = run + ex;
            p.win_cstart[g0 + c] = g0 + run + ex;
        }
        run += tot;
    }
    __syncthreads();
    // ---- stable placement: one warp walks the detections in order, lanes of the same component share a cursor bump
    if (tid < 32) {
        const unsigned lt = (1u << tid) - 1u;
        for (int base = 0; base < n; base += 32) {
            const int i = base + tid;
            const int r = i < n ? s_par[i] : -1;
            const unsigned m = __match_any_sync(0xffffffffu, r);
            const int leader = __ffs(m) - 1;
            int start = 0;
            if (tid == leader && r >= 0) {
                start = s_cnt[r];
                s_cnt[r] = start + __popc(m);
            }
            start = __shfl_sync(0xffffffffu, start, leader);
            if (i < n) s_par[i] = start + __popc(m & lt);
        }
    }
    __syncthreads();
    for (int i = tid; i < n; i += T) {
        const int q = s_par[i];
        p.comp_inv[g0 + i] = g0 + q;
        p.comp_perm[g0 + q] = g0 + i;
    }
    if (tid == 0) p.win_ncomp[g] = ncomp;
}

// ---- the same split for the solver, which has the first shortest-path tree of every graph by then (k_first_tree runs
//      before this kernel). All nodes of a tree branch - the nodes whose tree path starts with the same S arc - are
//      connected by tree arcs, which are arcs of the graph: the union-find starts with every detection hooked under its
//      branch root, and an arc only matters when the parents of its two ends differ - a few per cent of the arcs of a
//      tracking window, where the old kernel ran two finds and a compare-and-swap attempt for every arc (117 thread
//      instructions per arc, most of k_components' time). One thread per row: the common case is a load of the source
//      and one shared-memory read. Unreachable detections (no branch) are their own roots and are united through
//      their arcs like any other node, so the partition is exactly the connected components of the arc graph.
//      The kernel also moves the first tree from detection order (FtTmp) to the regrouped order the solver stages
//      from, turning parents and branch labels into positions on the way.
__global__ void __launch_bounds__(CC_TPB) k_components_tree(const CompParams p) {
    extern __shared__ int s_dyn[];
    __shared__ int s_scan[33];
    int *s_par = s_dyn;          // [cap] parent -> root -> component rank -> regrouped position
    int *s_cnt = s_dyn + p.cap;  // [cap] rank of a root -> detections per component -> placement cursor
    const int g = blockIdx.x + p.g_first;
    const int g0 = p.graph_ptr[g];
    const int n = p.graph_ptr[g + 1] - g0;
    const int T = blockDim.x, tid = threadIdx.x;
    if (n <= 0) {
        if (tid == 0) p.win_ncomp[g] = 0;
        return;
    }
    const int32_t *in_ptr = p.in_ptr + g0;
    const FtTmp *tmp = p.tree.tmp + g0;
    const bool overflow = (int64_t)in_ptr[n] > p.arc_capacity;  // arc buffer overflow: reported by the fill kernel
    const bool split = n <= p.cap && !overflow;
    if (!split) {  // kept whole: one component, identity order (an overflowed graph has no tree and is skipped later)
        for (int i = tid; i < n; i += T) {
            p.comp_perm[g0 + i] = g0 + i;
            p.comp_inv[g0 + i] = g0 + i;
            if (!overflow) p.tree.out[g0 + i] = tmp[i];
        }
        if (tid == 0) {
            p.win_ncomp[g] = 1;
            p.win_cstart[g0] = g0;
        }
        return;
    }
    for (int i = tid; i < n; i += T) {
        const int a = p.tree.anc[g0 + i];
        s_par[i] = a >= 0 ? a - g0 : i;  // (a branch root's label is the root itself)
    }
    __syncthreads();
    // ---- arcs between different branches. A warp takes 32 consecutive rows: their arcs are one contiguous range, read
    //      with coalesced, independent loads; the row of an arc is found by a binary search over the 32 row starts held
    //      by the lanes. The two ends of nearly every arc have the same parent (same branch, or united already): a
    //      find / unite only where they differ.
    {
        const int lane = lane_id(), nwarp = T >> 5;
        volatile int *vpar = s_par;
        for (int j0 = warp_id() * 32; j0 < n; j0 += nwarp * 32) {
            const int jl = j0 + lane;
            const int start = in_ptr[jl < n ? jl : n];  // rows beyond the graph: empty
            const int a0 = __shfl_sync(0xffffffffu, start, 0);
            const int a1 = in_ptr[j0 + 32 < n ? j0 + 32 : n];
            int src0 = a0 + lane < a1 ? p.in_src[a0 + lane] : 0;
            int src1 = a0 + 32 + lane < a1 ? p.in_src[a0 + 32 + lane] : 0;
            for (int eb = a0; eb < a1; eb += 32) {
                const int e = eb + lane;
                const int src2 = e + 64 < a1 ? p.in_src[e + 64] : 0;
                int lo = 0;  // the last row whose arcs start at or before e
#pragma unroll
                for (int step = 16; step > 0; step >>= 1) {
                    const int s = __shfl_sync(0xffffffffu, start, (lo + step) & 31);
                    if (s <= e) lo += step;
                }
                if (e < a1 && vpar[src0 - g0] != vpar[j0 + lo]) {
                    const int ri = cc_find(s_par, src0 - g0);
                    const int rj = cc_find(s_par, j0 + lo);
                    if (ri != rj) cc_unite(s_par, ri, rj);
                }
                src0 = src1;
                src1 = src2;
            }
        }
    }
    __syncthreads();
    // flatten: every detection points at its root (read-only chase; only the owner writes an entry)
    for (int i = tid; i < n; i += T) {
        volatile int *par = s_par;
        int r = par[i], q = par[r];
        while (q != r) {
            r = q;
            q = par[r];
        }
        par[i] = r;
    }
    __syncthreads();
    // ---- rank the roots (a branch root is the earliest detection of its branch and roots hook under smaller roots:
    //      the root of a component is its smallest detection, ascending root = ascending first detection)
    int ncomp = 0;
    for (int base = 0; base < n; base += T) {
        const int i = base + tid;
        const int f = (i < n && s_par[i] == i) ? 1 : 0;
        int tot;
        const int ex = block_excl_scan(f, s_scan, &tot);
        if (f) s_cnt[i] = ncomp + ex;
        ncomp += tot;
    }
    __syncthreads();
    for (int i = tid; i < n; i += T) s_par[i] = s_cnt[s_par[i]];  // own entry only: nobody else reads s_par[i] here
    __syncthreads();
    for (int c = tid; c < ncomp; c += T) s_cnt[c] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += T) atomicAdd(&s_cnt[s_par[i]], 1);
    __syncthreads();
    int run = 0;
    for (int base = 0; base < ncomp; base += T) {  // exclusive scan of the component sizes -> component starts
        const int c = base + tid;
        const int v = c < ncomp ? s_cnt[c] : 0;
        int tot;
        const int ex = block_excl_scan(v, s_scan, &tot);
        if (c < ncomp) {
            s_cnt[c] = run + ex;
            p.win_cstart[g0 + c] = g0 + run + ex;
        }
        run += tot;
    }
    __syncthreads();
    // ---- stable placement: one warp walks the detections in order, lanes of the same component share a cursor bump
    if (tid < 32) {
        const unsigned lt = (1u << tid) - 1u;
        for (int base = 0; base < n; base += 32) {
            const int i = base + tid;
            const int r = i < n ? s_par[i] : -1;
            const unsigned m = __match_any_sync(0xffffffffu, r);
            const int leader = __ffs(m) - 1;
            int start = 0;
            if (tid == leader && r >= 0) {
                start = s_cnt[r];
                s_cnt[r] = start + __popc(m);
            }
            start = __shfl_sync(0xffffffffu, start, leader);
            if (i < n) s_par[i] = start + __popc(m & lt);
        }
    }
    __syncthreads();
    // ---- permutation, and the first tree in the regrouped order (parents and branch labels become positions)
    for (int i = tid; i < n; i += T) {
        const int q = s_par[i];
        p.comp_inv[g0 + i] = g0 + q;
        p.comp_perm[g0 + q] = g0 + i;
        FtTmp t = tmp[i];
        t.par = t.par >= 0 ? g0 + s_par[t.par - g0] : t.par;
        t.anc = t.anc >= 0 ? g0 + s_par[t.anc - g0] : -1;
        p.tree.out[g0 + q] = t;  // one 32-byte record per position: one sector per detection
    }
    if (tid == 0) p.win_ncomp[g] = ncomp;
}

// component graphs of the whole batch: cgraph_ptr (positions in the regrouped order) and the graph each belongs to
__global__ void k_comp_offsets(int n_graphs, const int32_t *__restrict__ graph_ptr, const int32_t *__restrict__ win_ncomp,
                               int32_t *__restrict__ win_cbase, int32_t *__restrict__ n_cgraphs) {
    // single block: exclusive scan of the component counts
    __shared__ int s_scan[33];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int g0 = 0; g0 < n_graphs; g0 += blockDim.x) {
        const int g = g0 + threadIdx.x;
        const int v = g < n_graphs ? win_ncomp[g] : 0;
        int tot;
        const int ex = block_excl_scan(v, s_scan, &tot);
        if (g < n_graphs) win_cbase[g] = s_carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) s_carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        n_cgraphs[0] = s_carry;  // number of component graphs
        n_cgraphs[1] = 0;        // work-queue cursor of the solver
        n_cgraphs[2] = 0;
        n_cgraphs[3] = 0;
    }
}

__global__ void k_comp_fill(int n_graphs, const int32_t *__restrict__ graph_ptr, const int32_t *__restrict__ win_ncomp,
                            const int32_t *__restrict__ win_cbase, const int32_t *__restrict__ win_cstart,
                            int32_t *__restrict__ cgraph_ptr, int32_t *__restrict__ cgraph_win) {
    const int g = blockIdx.x;
    const int g0 = graph_ptr[g], nc = win_ncomp[g], cb = win_cbase[g];
    for (int c = threadIdx.x; c < nc; c += blockDim.x) {
        cgraph_ptr[cb + c] = win_cstart[g0 + c];
        cgraph_win[cb + c] = g;
    }
    if (g == n_graphs - 1 && threadIdx.x == 0) cgraph_ptr[cb + nc] = graph_ptr[n_graphs];
}

// max_graph_nodes sizes the shared memory of k_components (8 bytes per detection of the largest graph); the exported
// entry point does not know it and asks for the CC_MAX_SORT worst case.
// Graphs [g_first, g_first + g_count) are split in this call; `finish` numbers the components of ALL n_graphs graphs
// (every graph must have been through a call by then): a caller whose graphs arrive in pieces splits each piece as it
// lands and finishes once (mot3d_track_batch_host).
int32_t components_run(int32_t n_graphs, int64_t n_total, const int32_t *graph_ptr, int32_t max_graph_nodes,
                       const int32_t *in_ptr, const int32_t *in_src, int64_t arc_capacity, int32_t *comp_perm_out,
                       int32_t *comp_inv_out, int32_t *cgraph_ptr_out, int32_t *cgraph_graph_out,
                       int32_t *graph_cbase_out, int32_t *n_cgraphs_out, void *ws, size_t ws_bytes, void *stream,
                       int32_t g_first, int32_t g_count, bool finish, const FtOut *tree) {
    M3_CHECK_ARG(n_graphs >= 0 && n_total >= 0, "negative sizes");
    M3_CHECK_ARG(n_cgraphs_out, "n_cgraphs_out is NULL");
    cudaStream_t st = (cudaStream_t)stream;
    if (n_graphs == 0) {
        M3_CUDA(cudaMemsetAsync(n_cgraphs_out, 0, 4 * sizeof(int32_t), st));
        return MOT3D_OK;
    }
    M3_CHECK_ARG(graph_ptr && in_ptr && comp_perm_out && comp_inv_out && cgraph_ptr_out && cgraph_graph_out &&
                     graph_cbase_out && ws,
                 "NULL pointer");
    if (ws_bytes < mot3d_components_workspace_bytes(n_total, n_graphs)) {
        set_error("components workspace too small");
        return MOT3D_E_CAPACITY;
    }
    const size_t n = (size_t)(n_total > 0 ? n_total : 1), g = (size_t)n_graphs;
    char *w = (char *)ws;
    CompParams p;
    p.n_graphs = n_graphs;
    int64_t cap = max_graph_nodes;
    if (cap > n_total) cap = n_total;
    if (cap > CC_MAX_SORT) cap = CC_MAX_SORT;
    if (cap < 32) cap = 32;
    p.cap = (int32_t)cap;
    p.graph_ptr = graph_ptr;
    p.in_ptr = in_ptr;
    p.in_src = in_src;
    p.arc_capacity = arc_capacity;
    p.win_cstart = (int32_t *)w;
    w += (n * 4 + 255) & ~(size_t)255;
    p.win_ncomp = (int32_t *)w;
    w += (g * 4 + 255) & ~(size_t)255;
    int32_t *win_cbase = graph_cbase_out;
    p.comp_perm = comp_perm_out;
    p.comp_inv = comp_inv_out;
    const size_t smem = (size_t)cap * 2 * sizeof(int);
    if (smem > 48 * 1024)
        M3_CUDA(cudaFuncSetAttribute(k_components, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    p.g_first = g_first;
    if (g_count > 0 && tree) {  // the solver's path: the first tree is there, see k_components_tree
        p.tree = *tree;
        if (smem > 48 * 1024)
            M3_CUDA(cudaFuncSetAttribute(k_components_tree, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_components_tree<<<g_count, CC_TPB, smem, st>>>(p);
        M3_LAUNCH_CHECK("k_components_tree");
    } else if (g_count > 0) {
        k_components<<<g_count, CC_TPB, smem, st>>>(p);
        M3_LAUNCH_CHECK("k_components");
    }
    if (finish) {
        k_comp_offsets<<<1, 1024, 0, st>>>(n_graphs, graph_ptr, p.win_ncomp, win_cbase, n_cgraphs_out);
        M3_LAUNCH_CHECK("k_comp_offsets");
        k_comp_fill<<<n_graphs, 128, 0, st>>>(n_graphs, graph_ptr, p.win_ncomp, win_cbase, p.win_cstart, cgraph_ptr_out,
                                             cgraph_graph_out);
        M3_LAUNCH_CHECK("k_comp_fill");
    }
    return MOT3D_OK;
}

}  // namespace m3

using namespace m3;

extern "C" {

size_t mot3d_components_workspace_bytes(int64_t n_total, int32_t n_graphs) {
    size_t n = (size_t)(n_total > 0 ? n_total : 1), g = (size_t)(n_graphs > 0 ? n_graphs : 1);
    return ((n * 4 + 255) & ~(size_t)255) + ((g * 4 + 255) & ~(size_t)255) + 256;
}

int32_t mot3d_components(int32_t n_graphs, int64_t n_total, const int32_t *graph_ptr, const int32_t *in_ptr,
                         const int32_t *in_src, int64_t arc_capacity, int32_t *comp_perm_out, int32_t *comp_inv_out,
                         int32_t *cgraph_ptr_out, int32_t *cgraph_graph_out, int32_t *graph_cbase_out,
                         int32_t *n_cgraphs_out, void *ws,
                         size_t ws_bytes, void *stream) {
    return components_run(n_graphs, n_total, graph_ptr, CC_MAX_SORT, in_ptr, in_src, arc_capacity, comp_perm_out,
                          comp_inv_out, cgraph_ptr_out, cgraph_graph_out, graph_cbase_out, n_cgraphs_out, ws, ws_bytes,
                          stream, 0, n_graphs, true, nullptr);
}

}  // extern "C"
