// Shared device/host helpers for the mot3d_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mot3d_b200.h"

namespace m3 {

// ---------------------------------------------------------------------------------------- error plumbing
void set_error(const char *fmt, ...);
int32_t cuda_fail(cudaError_t e, const char *what);

#define M3_CHECK_ARG(cond, msg)                \
    do {                                       \
        if (!(cond)) {                         \
            m3::set_error("%s", msg);          \
            return MOT3D_E_INVALID;            \
        }                                      \
    } while (0)

#define M3_CUDA(call)                                        \
    do {                                                     \
        cudaError_t e__ = (call);                            \
        if (e__ != cudaSuccess) return m3::cuda_fail(e__, #call); \
    } while (0)

#define M3_LAUNCH_CHECK(name)                                     \
    do {                                                          \
        cudaError_t e__ = cudaGetLastError();                     \
        if (e__ != cudaSuccess) return m3::cuda_fail(e__, name);  \
    } while (0)

// What k_first_tree leaves per detection (in detection order; k_components_tree moves it to the regrouped order):
// exact distances of the pre / post node, parent of the pre node (detection index / PAR_*), branch label (detection
// whose S arc starts the node's tree path, -1 = unreachable) and the bit k_single_track's certificate asks for.
struct __align__(16) FtTmp {
    int64_t dpre, dpost;
    int32_t par, anc, cert, pad;
};
struct FtOut {  // first tree: detection order in, regrouped order out (parents and branch labels as positions)
    const FtTmp *tmp;
    const int32_t *anc;  // the branch labels once more, 4 bytes per detection: what the union-find starts from
    FtTmp *out;
};

// csrc/components.cu: mot3d_components with the shared-memory footprint sized by the largest graph of the batch
int32_t components_run(int32_t n_graphs, int64_t n_total, const int32_t *graph_ptr, int32_t max_graph_nodes,
                       const int32_t *in_ptr, const int32_t *in_src, int64_t arc_capacity, int32_t *comp_perm_out,
                       int32_t *comp_inv_out, int32_t *cgraph_ptr_out, int32_t *cgraph_graph_out,
                       int32_t *graph_cbase_out, int32_t *n_cgraphs_out, void *ws, size_t ws_bytes, void *stream,
                       int32_t g_first, int32_t g_count, bool finish, const FtOut *tree);

// csrc/solve.cu: mot3d_solve_batch in phases, for a caller whose graphs become ready in pieces (mot3d_track_batch_host):
// SOLVE_BEGIN once, SOLVE_FRONT for every range of graphs [g_first, g_first + g_count) whose in-CSR is complete
// (connected components + first shortest-path tree: per-graph kernels), SOLVE_BACK once when all of them are through
// (component numbering, single-track pass, work queue, successive shortest paths, first-path rule). Same arguments
// in every call.
enum { SOLVE_BEGIN = 1, SOLVE_FRONT = 2, SOLVE_BACK = 4, SOLVE_ALL = 7 };
int32_t solve_batch_phased(int phases, int32_t g_first, int32_t g_count, int32_t n_graphs, int64_t n_total,
                           const int32_t *graph_ptr, int32_t max_graph_nodes, const int32_t *tkey, int64_t arc_capacity,
                           const int32_t *in_ptr, const int32_t *in_src, const int64_t *in_cost,
                           const int64_t *entry_cost, const int64_t *node_cost, const int64_t *exit_cost,
                           int32_t *next_out, int32_t *prev_out, int32_t *n_paths_out, int64_t *total_cost_out,
                           int64_t *path_cost_out, int64_t *stats_out, void *ws, size_t ws_bytes, void *stream);

// Process-wide options (mot3d_set_option): test hooks and tuning knobs that used to be environment variables. An option
// that was never set (or was reset with NaN) reports false and leaves the built-in default alone.
enum Option {
    OPT_HOST_CHUNKS, OPT_HOST_SPLIT, OPT_HOST_EDGE_WEIGHT, OPT_HOST_TRACE, OPT_HOST_NO_PACK, OPT_HOST_PIECES,
    OPT_EDGE_CAP, OPT_EDGE_ACAP, OPT_SOLVE_NO_FAST, OPT_BIG_DELTA, OPT_COUNT
};
bool option(Option which, double *value);

constexpr int MOT3D_MAX_DEVICES = 64;  // per-device caches (launch parameters, host pipeline contexts)
constexpr int64_t INF = MOT3D_INF_COST;
constexpr int64_t INF_CUT = MOT3D_INF_COST / 2;  // anything >= this is "unreachable"

// ---------------------------------------------------------------------------------------- warp / block primitives
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }

__device__ __forceinline__ int warp_incl_scan(int v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane_id() >= d) v += t;
    }
    return v;
}

// Exclusive scan of one int per thread over the whole block; returns the exclusive prefix, *total gets the block sum.
// `scratch` = 33 ints of shared memory. Ends with a barrier, so it may be called back to back.
__device__ __forceinline__ int block_excl_scan(int v, int *scratch, int *total) {
    int incl = warp_incl_scan(v);
    int w = warp_id(), l = lane_id();
    int nw = (blockDim.x + 31) >> 5;
    if (l == 31) scratch[w] = incl;
    __syncthreads();
    if (w == 0) {
        int s = (l < nw) ? scratch[l] : 0;
        int si = warp_incl_scan(s);
        scratch[l] = si - s;
        if (l == 31) scratch[32] = si;
    }
    __syncthreads();
    int base = scratch[w];
    *total = scratch[32];
    __syncthreads();
    return base + incl - v;
}

struct KeyIdx {
    int64_t key;
    int idx;
};
__device__ __forceinline__ KeyIdx min_ki(KeyIdx a, KeyIdx b) {
    // ties -> smaller index, so results do not depend on the reduction shape
    return (b.key < a.key || (b.key == a.key && b.idx < a.idx)) ? b : a;
}
__device__ __forceinline__ KeyIdx shfl_down_ki(KeyIdx a, int d, unsigned mask = 0xffffffffu) {
    KeyIdx r;
    r.key = __shfl_down_sync(mask, a.key, d);
    r.idx = __shfl_down_sync(mask, a.idx, d);
    return r;
}
__device__ __forceinline__ KeyIdx shfl_xor_ki(KeyIdx a, int d, unsigned mask = 0xffffffffu) {
    KeyIdx r;
    r.key = __shfl_xor_sync(mask, a.key, d);
    r.idx = __shfl_xor_sync(mask, a.idx, d);
    return r;
}

// ---------------------------------------------------------------------------------------- TMA bulk copy (1-D) + mbarrier
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra.uni WAIT_DONE;\n"
        "bra.uni WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy through the TMA unit (SASS: UBLKCP). dst, src and bytes must be multiples of 16.
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// Stage g[c0 .. c0+cnt) into shared memory so that element e sits at s[e - c0 + shift]; the 16-byte aligned body
// goes through the TMA unit (cp.async.bulk -> UBLKCP), the ragged head/tail through ordinary loads.
// Returns shift. *bulk_bytes (thread 0 only) accumulates the bytes the mbarrier has to expect.
template <typename T>
__device__ __forceinline__ int stage_slice(const T *__restrict__ g, T *s, int64_t c0, int cnt, uint64_t *bar,
                                           uint32_t *bulk_bytes) {
    constexpr int Q = 16 / (int)sizeof(T);
    const uintptr_t first = (uintptr_t)(g + c0) / sizeof(T);
    const int shift = (int)(first % Q);
    int head = (Q - shift) % Q;
    if (head > cnt) head = cnt;
    int body = ((cnt - head) / Q) * Q;
    int tail = cnt - head - body;
    if (threadIdx.x == 0 && body > 0) {
        tma_load_1d(s + shift + head, g + c0 + head, (uint32_t)(body * sizeof(T)), bar);
        *bulk_bytes += (uint32_t)(body * sizeof(T));
    }
    if ((int)threadIdx.x < head) s[shift + threadIdx.x] = g[c0 + threadIdx.x];
    if ((int)threadIdx.x < tail) s[shift + head + body + threadIdx.x] = g[c0 + head + body + threadIdx.x];
    return shift;
}

// ---------------------------------------------------------------------------------------- exact quantisation
// Nearest integer to w * 1e7 evaluated exactly (ties to even) = what "{:0.7f}".format(w) prints
// (mot3d/mot3d.py:224), without the double rounding of llrint(w * 1e7).
__device__ __forceinline__ int64_t quantise_1e7(double w) {
    const double S = 1e7;
    double p = __dmul_rn(w, S);
    double n = rint(p);
    double r = fma(w, S, -n);  // exact w*S - n up to one rounding of a tiny value
    if (r > 0.5)
        n += 1.0;
    else if (r < -0.5)
        n -= 1.0;
    else if (r == 0.5) {
        if (fmod(n, 2.0) != 0.0) n += 1.0;
    } else if (r == -0.5) {
        if (fmod(n, 2.0) != 0.0) n -= 1.0;
    }
    return (int64_t)n;
}

// ---------------------------------------------------------------------------------------- arc cost pieces
struct JumpTable {
    double v[MOT3D_MAX_JUMP];  // -exp(-(jump-1)*sigma_jump), host-evaluated (mot3d_cost_spec::neg_exp_jump)
};

// bbox_size_similarity, mot3d/distance_functions.py:36-45
__device__ __forceinline__ double bbox_similarity(const double *__restrict__ bb, int64_t n, int64_t a, int64_t b) {
    // size = (ymax-ymin, xmax-xmin); 1 - (|h1-h2|/(max(h1,h2)+1e-12) + |w1-w2|/(max(w1,w2)+1e-12))/2
    double h1 = bb[3 * n + a] - bb[1 * n + a];
    double w1 = bb[2 * n + a] - bb[0 * n + a];
    double h2 = bb[3 * n + b] - bb[1 * n + b];
    double w2 = bb[2 * n + b] - bb[0 * n + b];
    double th = fabs(h1 - h2) / (fmax(h1, h2) + 1e-12);
    double tw = fabs(w1 - w2) / (fmax(w1, w2) + 1e-12);
    return 1.0 - (th + tw) / 2.0;
}

// w = (-exp(-(jump-1)*sigma_jump)) * ((exp(-dist^2/sigma_d^2) [* w_hist]) * w_bbox)
__device__ __forceinline__ double distance_factor(double s, double sigma_distance_sq) {
    double dist = sqrt(s);
    double q = dist * dist;  // the reference squares the rounded distance (dist**2)
    return exp(-q / sigma_distance_sq);
}

}  // namespace m3
