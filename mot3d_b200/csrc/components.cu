// Connected components of every graph of a batch, and the regrouping of its detections by component.
//
// Why: a tracking window decomposes into groups of detections that never come within max_distance/max_jump of each
// other; the min-cost flow of the window is the union of the min-cost flows of the groups (the arcs through S and T
// are the only coupling and the acceptance rule "keep a path while its cost < 0", wrapper.cpp:263-267, applies to
// every path on its own). Solving the groups as independent small graphs turns ~100 dependent augmentations per
// window into ~60 independent chains of 1-3, and the per-augmentation scans shrink from the window to the group.
// The reference has no counterpart (muSSP handles one graph sequentially); results are identical.
//
// One block per graph: min-label propagation with pointer jumping over the transition arcs (labels in global memory),
// then a bitonic sort of (component rank, detection) keys in shared memory gives the detections grouped by component,
// ascending detection index inside a component (= frame order, which the solver's layer sweeps need). Graphs larger
// than CC_MAX_SORT detections are kept whole.
#include "common.cuh"

namespace m3 {

constexpr int CC_TPB = 256;
constexpr int CC_MAX_SORT = 8192;

struct CompParams {
    int32_t n_graphs;
    const int32_t *graph_ptr;
    const int32_t *in_ptr;
    const int32_t *in_src;
    int64_t arc_capacity;
    int32_t *label;        // [n_total] scratch: component root (local index) per detection
    int32_t *comp_perm;    // [n_total] out: detection (global index) at every regrouped position
    int32_t *comp_inv;     // [n_total] out: regrouped position (global) of every detection
    int32_t *win_ncomp;    // [n_graphs] out: components per graph
    int32_t *win_cstart;   // [n_total] out: graph g's component c starts at regrouped position win_cstart[g0 + c]
};

__global__ void __launch_bounds__(CC_TPB) k_components(const CompParams p) {
    __shared__ int s_keys[CC_MAX_SORT];
    __shared__ int s_scan[33];
    __shared__ int s_changed;
    const int g = blockIdx.x;
    const int g0 = p.graph_ptr[g];
    const int n = p.graph_ptr[g + 1] - g0;
    const int T = blockDim.x, tid = threadIdx.x;
    if (n <= 0) {
        if (tid == 0) p.win_ncomp[g] = 0;
        return;
    }
    int32_t *label = p.label + g0;
    const int32_t *in_ptr = p.in_ptr + g0;
    const bool overflow = (int64_t)in_ptr[n] > p.arc_capacity;  // arc buffer overflow: reported by the fill kernel
    const bool split = n <= CC_MAX_SORT && !overflow;
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
    for (int i = tid; i < n; i += T) label[i] = i;
    __syncthreads();
    // ---- min-label propagation over the arcs + pointer jumping, until nothing changes
    for (int iter = 0; iter < n + 2; iter++) {
        if (tid == 0) s_changed = 0;
        __syncthreads();
        int ch = 0;
        for (int j = tid; j < n; j += T) {
            const int e1 = in_ptr[j + 1];
            int lj = label[j];
            for (int e = in_ptr[j]; e < e1; e++) {
                const int i = p.in_src[e] - g0;
                const int li = label[i];
                if (li < lj) {
                    lj = li;
                } else if (lj < li) {
                    atomicMin(&label[i], lj);
                    ch = 1;
                }
            }
            if (lj < label[j]) {
                atomicMin(&label[j], lj);
                ch = 1;
            }
        }
        __syncthreads();
        for (int i = tid; i < n; i += T) {  // jump to the label's label until a fixed point
            int l = label[i];
            int ll = label[l];
            while (ll != l) {
                l = ll;
                ll = label[l];
            }
            if (l != label[i]) {
                label[i] = l;
                ch = 1;
            }
        }
        if (ch) s_changed = 1;
        __syncthreads();
        if (!s_changed) break;
        __syncthreads();
    }
    // ---- rank the roots (ascending root index) and sort (rank, detection) keys
    int32_t *root_rank = p.win_cstart + g0;  // scratch until the final loop below rewrites it
    int ncomp = 0;
    for (int base = 0; base < n; base += T) {
        const int i = base + tid;
        const int f = (i < n && label[i] == i) ? 1 : 0;
        int tot;
        const int ex = block_excl_scan(f, s_scan, &tot);
        if (f) root_rank[i] = ncomp + ex;
        ncomp += tot;
    }
    __syncthreads();
    int npow = 1;
    while (npow < n) npow <<= 1;
    for (int i = tid; i < npow; i += T)  // key = rank << 13 | detection (n <= 8192); padding sorts last
        s_keys[i] = (i < n) ? ((root_rank[label[i]] << 13) | i) : 0x7fffffff;
    __syncthreads();
    for (int k = 2; k <= npow; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < npow; i += T) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const int a = s_keys[i], b = s_keys[ixj];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) {
                        s_keys[i] = b;
                        s_keys[ixj] = a;
                    }
                }
            }
            __syncthreads();
        }
    }
    for (int q = tid; q < n; q += T) {
        const int key = s_keys[q];
        const int i = key & 8191;
        p.comp_perm[g0 + q] = g0 + i;
        p.comp_inv[g0 + i] = g0 + q;
        const int rank = key >> 13;
        if (q == 0 || (s_keys[q - 1] >> 13) != rank) p.win_cstart[g0 + rank] = g0 + q;
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

}  // namespace m3

using namespace m3;

extern "C" {

size_t mot3d_components_workspace_bytes(int64_t n_total, int32_t n_graphs) {
    size_t n = (size_t)(n_total > 0 ? n_total : 1), g = (size_t)(n_graphs > 0 ? n_graphs : 1);
    return ((n * 4 + 255) & ~(size_t)255) * 2 + ((g * 4 + 255) & ~(size_t)255) + 256;
}

int32_t mot3d_components(int32_t n_graphs, int64_t n_total, const int32_t *graph_ptr, const int32_t *in_ptr,
                         const int32_t *in_src, int64_t arc_capacity, int32_t *comp_perm_out, int32_t *comp_inv_out,
                         int32_t *cgraph_ptr_out, int32_t *cgraph_graph_out, int32_t *graph_cbase_out,
                         int32_t *n_cgraphs_out, void *ws,
                         size_t ws_bytes, void *stream) {
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
    p.graph_ptr = graph_ptr;
    p.in_ptr = in_ptr;
    p.in_src = in_src;
    p.arc_capacity = arc_capacity;
    p.label = (int32_t *)w;
    w += (n * 4 + 255) & ~(size_t)255;
    p.win_cstart = (int32_t *)w;
    w += (n * 4 + 255) & ~(size_t)255;
    p.win_ncomp = (int32_t *)w;
    w += (g * 4 + 255) & ~(size_t)255;
    int32_t *win_cbase = graph_cbase_out;
    p.comp_perm = comp_perm_out;
    p.comp_inv = comp_inv_out;
    k_components<<<n_graphs, CC_TPB, 0, st>>>(p);
    M3_LAUNCH_CHECK("k_components");
    k_comp_offsets<<<1, 1024, 0, st>>>(n_graphs, graph_ptr, p.win_ncomp, win_cbase, n_cgraphs_out);
    M3_LAUNCH_CHECK("k_comp_offsets");
    k_comp_fill<<<n_graphs, 128, 0, st>>>(n_graphs, graph_ptr, p.win_ncomp, win_cbase, p.win_cstart, cgraph_ptr_out,
                                         cgraph_graph_out);
    M3_LAUNCH_CHECK("k_comp_fill");
    return MOT3D_OK;
}

}  // extern "C"
