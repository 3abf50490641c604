// C-ABI plumbing: version, thread-local error text, and the one-call host-buffer entry point
// (pack -> H2D -> keys -> frame sort -> single-pass arc build -> solve -> tracks -> D2H).
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>

#include "common.cuh"

namespace m3 {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int32_t cuda_fail(cudaError_t e, const char *what) {
    set_error("CUDA error in %s: %s", what, cudaGetErrorString(e));
    return MOT3D_E_CUDA;
}

// bump allocator over the caller's device workspace
struct Arena {
    char *base;
    size_t size, used;
    bool ok;
    Arena(void *p, size_t n) : base((char *)p), size(n), used(0), ok(true) {}
    template <typename T>
    T *take(int64_t count) {
        size_t bytes = ((size_t)(count > 0 ? count : 1) * sizeof(T) + 255) & ~(size_t)255;
        if (used + bytes > size) {
            ok = false;
            return nullptr;
        }
        T *r = (T *)(base + used);
        used += bytes;
        return r;
    }
};

static size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }

}  // namespace m3

using namespace m3;

extern "C" {

const char *mot3d_version(void) { return "mot3d_b200 0.1 (sm_100a)"; }
const char *mot3d_last_error(void) { return g_err; }

}  // extern "C"

namespace m3 {
// Only K_g + 1 entries of graph g's slice of track_ptr are meaningful (include/mot3d_b200.h): they are packed for the
// D2H copy (0.5 MB instead of 24 MB per 6 M detections) and spread out again on the host.
__global__ void k_track_ptr_offsets(int n_graphs, const int32_t *__restrict__ track_count, int32_t *__restrict__ off) {
    // single block: off[g] = sum_{h<g} (K_h + 1)
    __shared__ int s_scan[33];
    int run = 0;
    for (int g0 = 0; g0 < n_graphs; g0 += blockDim.x) {
        const int g = g0 + threadIdx.x;
        const int v = g < n_graphs ? track_count[g] + 1 : 0;
        int tot;
        const int ex = block_excl_scan(v, s_scan, &tot);
        if (g < n_graphs) off[g] = run + ex;
        run += tot;
    }
}
__global__ void k_track_ptr_pack(int n_graphs, const int32_t *__restrict__ graph_ptr,
                                 const int32_t *__restrict__ track_count, const int32_t *__restrict__ off,
                                 const int32_t *__restrict__ track_ptr, int32_t *__restrict__ packed) {
    const int g = blockIdx.x;
    if (g >= n_graphs) return;
    const int32_t *src = track_ptr + graph_ptr[g] + g;
    int32_t *dst = packed + off[g];
    for (int k = threadIdx.x; k <= track_count[g]; k += blockDim.x) dst[k] = src[k];
}
}  // namespace m3

namespace m3 {
int32_t launch_make_keys_based(const int32_t *frame, const int32_t *graph_ptr, int32_t n_graphs, const int64_t *base,
                               int32_t *tkey, int32_t *err, cudaStream_t st);
int32_t launch_frame_sort_range(const mot3d_dets *d, int64_t row_begin, int64_t row_end, cudaStream_t st);
}  // namespace m3

// device bytes one pipeline pass over (n, n_graphs, edge_capacity) needs
static size_t pass_bytes(int64_t n, int32_t n_graphs, int64_t edge_capacity) {
    size_t b = 0;
    int64_t n1 = n > 0 ? n : 1, g1 = n_graphs > 0 ? n_graphs : 1, e1 = edge_capacity > 0 ? edge_capacity : 1;
    b += align256((size_t)(g1 + 1) * 4);        // graph_ptr
    b += align256((size_t)n1 * 4) * 2;          // frame, tkey
    b += align256((size_t)n1 * 8 * 3);          // pos
    b += align256((size_t)n1 * 8);              // conf
    b += align256((size_t)n1);                  // flags
    b += align256((size_t)n1 * 8 * 4);          // bbox
    b += align256((size_t)g1 * 8);              // graph base
    b += align256((size_t)n1 * 8 * 3) + align256((size_t)n1 * 4);  // x-sorted copies + permutation
    b += align256((size_t)n1 * 8) * 3;          // entry/node/exit cost
    b += align256((size_t)(n1 + 1) * 4);        // row_ptr
    b += align256(mot3d_edge_build_workspace_bytes(n1));
    b += align256((size_t)e1 * 4);              // col
    b += align256((size_t)e1 * 8);              // cost
    b += align256(4);                           // overflow flag
    b += align256((size_t)n1 * 4) * 2;          // next, prev
    b += align256((size_t)g1 * 4);              // n_paths
    b += align256((size_t)g1 * 8);              // total
    b += align256(mot3d_solve_workspace_bytes(n1, g1, e1));
    b += align256((size_t)g1 * 4);              // track_count
    b += align256((size_t)(n1 + g1) * 4);       // track_ptr
    b += align256((size_t)n1 * 4);              // track_det
    b += align256((size_t)g1 * 4) + align256((size_t)(n1 + g1) * 4);  // packed track_ptr + offsets
    return b;
}

// ---- software pipeline of mot3d_track_batch_host -----------------------------------------------------------------
// Whole graphs are independent, so a large batch is cut into chunks of consecutive graphs. Every chunk runs its
// H2D copies, its 18 kernels and its D2H copies on a stream of its own over its own slice of the workspace: the copy
// engines move chunk c+1 in and chunk c-1 out while the SMs work on chunk c, and the tail of one chunk's solver
// (a few long components) overlaps the next chunk's build.
constexpr int HOST_MAX_CHUNKS = 8;
constexpr int HOST_MAX_PIECES = 8;               // H2D + build pieces per chunk
constexpr int64_t HOST_PIECE_MIN_DETS = 400000;
constexpr int64_t HOST_CHUNK_MIN_DETS = 400000;  // below this a chunk cannot fill the GPU

struct HostCtx {
    int dev = -1;
    cudaStream_t s[HOST_MAX_CHUNKS];  // kernels + D2H of chunk c; earlier chunks have the higher priority
    cudaStream_t copy_in;             // all H2D copies, in chunk order
    cudaEvent_t start, in[HOST_MAX_CHUNKS * HOST_MAX_PIECES], done[HOST_MAX_CHUNKS];
    int64_t *pinned_gbase = nullptr;  // key bases per graph (H2D source), grown with pinned_gptr
    int32_t *pinned = nullptr;  // [2 * HOST_MAX_CHUNKS]: overflow flag, arc count per chunk
    int32_t *pinned_gptr = nullptr;  // chunk-local graph_ptr arrays (H2D source), grown on demand
    size_t gptr_cap = 0;
    cudaEvent_t t0 = nullptr, trace[HOST_MAX_CHUNKS][4];  // "host_trace": H2D done, build done, solve done, D2H done
    // released by mot3d_host_release (not from a destructor: at process exit the CUDA runtime may be gone already)
    void release() {
        if (dev < 0) return;
        int cur = -1;
        const bool switched = cudaGetDevice(&cur) == cudaSuccess && cur != dev && cudaSetDevice(dev) == cudaSuccess;
        for (int i = 0; i < HOST_MAX_CHUNKS; i++) {
            cudaStreamDestroy(s[i]);
            cudaEventDestroy(done[i]);
        }
        for (int i = 0; i < HOST_MAX_CHUNKS * HOST_MAX_PIECES; i++) cudaEventDestroy(in[i]);
        cudaStreamDestroy(copy_in);
        cudaEventDestroy(start);
        if (t0) {
            cudaEventDestroy(t0);
            for (int c = 0; c < HOST_MAX_CHUNKS; c++)
                for (int k = 0; k < 4; k++) cudaEventDestroy(trace[c][k]);
            t0 = nullptr;
        }
        if (pinned) cudaFreeHost(pinned);
        if (pinned_gptr) cudaFreeHost(pinned_gptr);
        if (pinned_gbase) cudaFreeHost(pinned_gbase);
        pinned = pinned_gptr = nullptr;
        pinned_gbase = nullptr;
        gptr_cap = 0;
        dev = -1;
        if (switched) cudaSetDevice(cur);
        cudaGetLastError();
    }
};
struct HostCtxSet {
    HostCtx ctx[MOT3D_MAX_DEVICES];
};
static HostCtxSet *host_ctx_set() {
    static thread_local HostCtxSet set;
    return &set;
}

// per host thread and device: streams, events and a pinned scratch for the two scalars read back per chunk
static HostCtx *host_ctx() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MOT3D_MAX_DEVICES) return nullptr;
    HostCtx *c = &host_ctx_set()->ctx[dev];
    if (c->dev == dev) return c;
    int prio_lo = 0, prio_hi = 0;  // numerically lower = higher priority
    if (cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi) != cudaSuccess) return nullptr;
    for (int i = 0; i < HOST_MAX_CHUNKS; i++) {
        int prio = prio_hi + i;
        if (prio > prio_lo) prio = prio_lo;
        if (cudaStreamCreateWithPriority(&c->s[i], cudaStreamNonBlocking, prio) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&c->done[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
    }
    for (int i = 0; i < HOST_MAX_CHUNKS * HOST_MAX_PIECES; i++)
        if (cudaEventCreateWithFlags(&c->in[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaStreamCreateWithPriority(&c->copy_in, cudaStreamNonBlocking, prio_hi) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&c->start, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaHostAlloc((void **)&c->pinned, sizeof(int32_t) * 2 * HOST_MAX_CHUNKS, cudaHostAllocDefault) != cudaSuccess)
        return nullptr;
    c->dev = dev;
    return c;
}

// graphs [ga, gb) of the batch = detections [d0, d1): enqueue H2D, kernels and D2H on `st`. h_gptr = chunk-local
// graph_ptr, h_gbase = key bases of its graphs (pinned host memory, must stay alive until the stream has consumed them);
// h_flags[0..1] receive overflow flag, arc count. The inputs arrive in `pieces` slices of whole graphs (copy stream
// `cin`, one event each): keys, node costs, frame sort and the single-pass arc build of a piece run while the next one
// is on its way - the arc build continues the CSR of the piece before it through a device-side counter - and only the
// solver waits for the whole chunk.
static int32_t enqueue_chunk(const mot3d_host_batch *hb, const mot3d_cost_spec *spec, mot3d_host_result *res, Arena &A,
                             int32_t ga, int32_t gb, const int32_t *h_gptr, const int64_t *h_gbase, int32_t max_nodes,
                             int64_t edge_capacity, int32_t *h_flags, cudaStream_t st, cudaStream_t cin,
                             cudaEvent_t *in_done, int pieces, cudaEvent_t *trace, cudaEvent_t counts_done,
                             int32_t **d_packed_out) {
    const int64_t d0 = hb->graph_ptr[ga], d1 = hb->graph_ptr[gb];
    const int64_t n = d1 - d0, N = hb->n;
    const int32_t G = gb - ga;
    int32_t *d_gptr = A.take<int32_t>(G + 1);
    int32_t *d_frame = A.take<int32_t>(n);
    int32_t *d_tkey = A.take<int32_t>(n);
    double *d_pos = A.take<double>(n * hb->dim);
    double *d_conf = A.take<double>(n);
    uint8_t *d_flags = hb->flags ? A.take<uint8_t>(n) : nullptr;
    double *d_bbox = (spec->use_bbox) ? A.take<double>(n * 4) : nullptr;
    int64_t *d_gbase = A.take<int64_t>(G);
    double *d_spos = A.take<double>(n * hb->dim);
    int32_t *d_sperm = A.take<int32_t>(n);
    int64_t *d_en = A.take<int64_t>(n);
    int64_t *d_cn = A.take<int64_t>(n);
    int64_t *d_ex = A.take<int64_t>(n);
    int32_t *d_rowptr = A.take<int32_t>(n + 1);
    size_t build_b = mot3d_edge_build_workspace_bytes(n);
    char *d_build = A.take<char>((int64_t)build_b);
    int32_t *d_col = A.take<int32_t>(edge_capacity);
    int64_t *d_cost = A.take<int64_t>(edge_capacity);
    int32_t *d_over = A.take<int32_t>(1);
    int32_t *d_next = A.take<int32_t>(n);
    int32_t *d_prev = A.take<int32_t>(n);
    int32_t *d_np = A.take<int32_t>(G);
    int64_t *d_total = A.take<int64_t>(G);
    size_t solve_b = mot3d_solve_workspace_bytes(n, G, edge_capacity);
    char *d_solve = A.take<char>((int64_t)solve_b);
    int32_t *d_tcount = A.take<int32_t>(G);
    int32_t *d_tptr = A.take<int32_t>(n + G);
    int32_t *d_tdet = A.take<int32_t>(n);
    if (!A.ok) return MOT3D_E_CAPACITY;  // the caller retries with fewer chunks / reports

    int64_t *d_arcs = A.take<int64_t>(1);
    int32_t *d_poff = d_packed_out ? A.take<int32_t>(G) : nullptr;
    int32_t *d_packed = d_packed_out ? A.take<int32_t>(n + G) : nullptr;
    if (!A.ok) return MOT3D_E_CAPACITY;
    M3_CUDA(cudaMemcpyAsync(d_gptr, h_gptr, (size_t)(G + 1) * 4, cudaMemcpyHostToDevice, cin));
    M3_CUDA(cudaMemcpyAsync(d_gbase, h_gbase, (size_t)G * 8, cudaMemcpyHostToDevice, cin));
    M3_CUDA(cudaMemsetAsync(d_over, 0, 4, st));
    M3_CUDA(cudaMemsetAsync(d_rowptr, 0, 4, st));
    mot3d_dets d;
    memset(&d, 0, sizeof(d));
    d.n = n;
    d.dim = hb->dim;
    d.tkey = d_tkey;
    d.pos = d_pos;
    d.conf = d_conf;
    d.flags = d_flags;
    d.bbox = d_bbox;
    d.sorted_pos = d_spos;
    d.sorted_perm = d_sperm;
    int32_t rc;
    if (pieces < 1) pieces = 1;
    int32_t g_lo = 0, g_front = 0;
    bool first = true;
    // The solve runs in phases (solve_batch_phased): its per-graph front end - connected components and first
    // shortest-path tree, a third of the stage - follows the arc build of every piece while the next piece is still on
    // the bus; only the component queue and the successive shortest paths wait for the whole chunk.
    auto solve_phase = [&](int phases, int32_t gf, int32_t gc) {
        return solve_batch_phased(phases, gf, gc, G, n, d_gptr, max_nodes, d_tkey, edge_capacity, d_rowptr, d_col, d_cost,
                                  d_en, d_cn, d_ex, d_next, d_prev, d_np, d_total, nullptr, nullptr, d_solve, solve_b, st);
    };
    if ((rc = solve_phase(SOLVE_BEGIN, 0, 0))) return rc;
    for (int pc = 0; pc < pieces; pc++) {
        // graphs [g_lo, g_hi) of the chunk: cut where the detection count passes the next multiple of n / pieces
        int32_t g_hi = g_lo;
        const int64_t target = pc + 1 == pieces ? n : n * (pc + 1) / pieces;
        while (g_hi < G && h_gptr[g_hi] < target) g_hi++;
        if (pc + 1 == pieces) g_hi = G;
        if (g_hi == g_lo) continue;
        const int64_t r0 = h_gptr[g_lo], r1 = h_gptr[g_hi], m = r1 - r0;
        if (m > 0) {
            M3_CUDA(cudaMemcpyAsync(d_frame + r0, hb->frame + d0 + r0, (size_t)m * 4, cudaMemcpyHostToDevice, cin));
            for (int k = 0; k < hb->dim; k++)
                M3_CUDA(cudaMemcpyAsync(d_pos + (size_t)k * n + r0, hb->pos + (size_t)k * N + d0 + r0, (size_t)m * 8,
                                        cudaMemcpyHostToDevice, cin));
            M3_CUDA(cudaMemcpyAsync(d_conf + r0, hb->conf + d0 + r0, (size_t)m * 8, cudaMemcpyHostToDevice, cin));
            if (d_flags) M3_CUDA(cudaMemcpyAsync(d_flags + r0, hb->flags + d0 + r0, (size_t)m, cudaMemcpyHostToDevice, cin));
            if (d_bbox)
                for (int k = 0; k < 4; k++)
                    M3_CUDA(cudaMemcpyAsync(d_bbox + (size_t)k * n + r0, hb->bbox + (size_t)k * N + d0 + r0, (size_t)m * 8,
                                            cudaMemcpyHostToDevice, cin));
        }
        if (cin != st) {  // inputs travel on the copy stream (FIFO over chunks and pieces); the chunk's stream picks up
            M3_CUDA(cudaEventRecord(in_done[pc], cin));
            M3_CUDA(cudaStreamWaitEvent(st, in_done[pc], 0));
        }
        if (m > 0) {
            if ((rc = launch_make_keys_based(d_frame, d_gptr + g_lo, g_hi - g_lo, d_gbase + g_lo, d_tkey, d_over, st))) return rc;
            mot3d_dets dp = d;  // per-detection outputs of this piece: plain offsets
            dp.n = m;
            dp.conf = d_conf + r0;
            dp.flags = d_flags ? d_flags + r0 : nullptr;
            if ((rc = mot3d_node_costs(&dp, spec, d_en + r0, d_cn + r0, d_ex + r0, st))) return rc;
            if ((rc = launch_frame_sort_range(&d, r0, r1, st))) return rc;
            if ((rc = mot3d_edge_build(&d, spec, MOT3D_DIR_IN, r0, r1, d_rowptr, edge_capacity, d_col, d_cost, nullptr,
                                       d_over, first ? nullptr : d_arcs, d_arcs, d_build, build_b, st)))
                return rc;
            first = false;
        }
        // (a launch of the per-graph kernels lasts as long as one graph takes, however few graphs it has: the pieces
        // are collected into two launches, one when half of the graphs are built - hidden behind the copies of the other
        // half - and one at the end; measured: 2 launches 6.97 ms per call, 3 uneven ones 7.27, 4 7.38, none 7.40)
        if (g_hi - g_front >= (G + 1) / 2) {
            if ((rc = solve_phase(SOLVE_FRONT, g_front, g_hi - g_front))) return rc;
            g_front = g_hi;
        }
        if (trace && pc == 0) M3_CUDA(cudaEventRecord(trace[0], st));
        g_lo = g_hi;
    }
    if (trace) M3_CUDA(cudaEventRecord(trace[1], st));
    if (g_front < G && (rc = solve_phase(SOLVE_FRONT, g_front, G - g_front))) return rc;
    if ((rc = solve_phase(SOLVE_BACK, 0, 0))) return rc;
    if ((rc = mot3d_extract_tracks(G, d_gptr, d_next, d_prev, d_tcount, d_tptr, d_tdet, st))) return rc;
    if (trace) M3_CUDA(cudaEventRecord(trace[2], st));

    // the per-graph slices of the outputs sit at graph_ptr[g] + g / graph_ptr[g] (include/mot3d_b200.h), so a chunk's
    // outputs are one contiguous slice of the caller's arrays
    M3_CUDA(cudaMemcpyAsync(res->track_count + ga, d_tcount, (size_t)G * 4, cudaMemcpyDeviceToHost, st));
    if (d_packed_out) {
        // the caller copies the packed track_ptr entries once it has the counts (they arrive first) and spreads them
        M3_CUDA(cudaEventRecord(counts_done, st));
        k_track_ptr_offsets<<<1, 1024, 0, st>>>(G, d_tcount, d_poff);
        M3_LAUNCH_CHECK("k_track_ptr_offsets");
        k_track_ptr_pack<<<G, 128, 0, st>>>(G, d_gptr, d_tcount, d_poff, d_tptr, d_packed);
        M3_LAUNCH_CHECK("k_track_ptr_pack");
        *d_packed_out = d_packed;
    } else {
        M3_CUDA(cudaMemcpyAsync(res->track_ptr + d0 + ga, d_tptr, (size_t)(n + G) * 4, cudaMemcpyDeviceToHost, st));
    }
    M3_CUDA(cudaMemcpyAsync(res->track_det + d0, d_tdet, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    M3_CUDA(cudaMemcpyAsync(res->total_cost + ga, d_total, (size_t)G * 8, cudaMemcpyDeviceToHost, st));
    M3_CUDA(cudaMemcpyAsync(h_flags, d_over, 4, cudaMemcpyDeviceToHost, st));
    M3_CUDA(cudaMemcpyAsync(h_flags + 1, d_rowptr + n, 4, cudaMemcpyDeviceToHost, st));
    if (trace) M3_CUDA(cudaEventRecord(trace[3], st));
    return MOT3D_OK;
}

namespace m3 {
static const char *const kOptionNames[OPT_COUNT] = {"host_chunks", "host_split", "host_edge_weight", "host_trace",
                                                    "host_no_pack", "host_pieces", "edge_cap", "edge_acap",
                                                    "solve_no_fast", "big_delta"};
static std::atomic<uint64_t> g_option_bits[OPT_COUNT];  // the double's bits; 0 = unset (0.0 is stored as -0.0's twin)
static std::atomic<uint32_t> g_option_set{0};
bool option(Option which, double *value) {
    if (!((g_option_set.load(std::memory_order_acquire) >> which) & 1u)) return false;
    const uint64_t bits = g_option_bits[which].load(std::memory_order_relaxed);
    memcpy(value, &bits, sizeof(double));
    return true;
}
}  // namespace m3

extern "C" {

int32_t mot3d_host_release(void) {
    HostCtxSet *set = host_ctx_set();
    for (int d = 0; d < MOT3D_MAX_DEVICES; d++) set->ctx[d].release();
    return MOT3D_OK;
}

int32_t mot3d_set_option(const char *name, double value) {
    M3_CHECK_ARG(name, "option name is NULL");
    for (int k = 0; k < OPT_COUNT; k++)
        if (strcmp(name, kOptionNames[k]) == 0) {
            if (value != value) {  // NaN: back to the built-in default
                g_option_set.fetch_and(~(1u << k), std::memory_order_release);
            } else {
                uint64_t bits;
                memcpy(&bits, &value, sizeof(double));
                g_option_bits[k].store(bits, std::memory_order_relaxed);
                g_option_set.fetch_or(1u << k, std::memory_order_release);
            }
            return MOT3D_OK;
        }
    set_error("unknown option '%s'", name);
    return MOT3D_E_INVALID;
}

size_t mot3d_track_batch_workspace_bytes(int64_t n, int32_t n_graphs, int64_t edge_capacity) {
    if (n < 0 || n_graphs < 0 || edge_capacity < 0) return 0;
    // every array is linear in (n, n_graphs, edge_capacity) up to its 256-byte alignment and a few constants, so the
    // chunks of the software pipeline fit into the single-pass size plus a fixed allowance per chunk
    return pass_bytes(n, n_graphs, edge_capacity) + (size_t)HOST_MAX_CHUNKS * 65536 + 4096;
}

int32_t mot3d_track_batch_host(const mot3d_host_batch *hb, const mot3d_cost_spec *spec, mot3d_host_result *res,
                               void *device_ws, size_t ws_bytes, int64_t edge_capacity, void *stream) {
    M3_CHECK_ARG(hb && spec && res, "NULL batch/spec/result");
    M3_CHECK_ARG(hb->n_graphs >= 0 && hb->n >= 0, "negative sizes");
    M3_CHECK_ARG(!spec->use_hist, "mot3d_track_batch_host: colour histograms go through the device entry points");
    res->n_edges = 0;
    res->h2d_bytes = res->d2h_bytes = 0;
    if (hb->n == 0 || hb->n_graphs == 0) return MOT3D_OK;
    M3_CHECK_ARG(hb->graph_ptr && hb->frame && hb->pos && hb->conf, "NULL host input");
    M3_CHECK_ARG(!spec->use_bbox || hb->bbox, "use_bbox set but bbox is NULL");
    M3_CHECK_ARG(res->track_count && res->track_ptr && res->track_det && res->total_cost, "NULL host output");
    M3_CHECK_ARG(device_ws, "device_ws is NULL");
    M3_CHECK_ARG(edge_capacity >= 0, "negative edge capacity");
    cudaStream_t user = (cudaStream_t)stream;
    const int64_t n = hb->n;
    const int32_t G = hb->n_graphs;
    for (int g = 0; g < G; g++)
        M3_CHECK_ARG(hb->graph_ptr[g + 1] >= hb->graph_ptr[g], "graph_ptr must be non-decreasing");
    M3_CHECK_ARG(hb->graph_ptr[0] == 0 && hb->graph_ptr[G] == n, "graph_ptr must span [0, n]");

    // Number of chunks: MOT3D_HOST_CHUNKS, else one. Every solver launch has a floor of ~2.4 ms whatever its share of
    // the batch (front end + its longest component, which is sequential work; profiles/r01g_summary.md), so two
    // launches cost more than the copy overlap gains until a batch is several times the size of the bench's 6 M
    // detections: with one chunk the pieces hide keys / sort / build behind the H2D copies (8.0 ms per 6 M detections;
    // 2 chunks x 8 pieces 8.5 ms, 4 chunks 10.4 ms).
    int want = n >= 40 * HOST_CHUNK_MIN_DETS ? 2 : 1;
    double ov;
    if (option(OPT_HOST_CHUNKS, &ov)) want = (int)ov;
    if (want > HOST_MAX_CHUNKS) want = HOST_MAX_CHUNKS;
    if (want > G) want = G;
    if (want < 1) want = 1;
    double first_share = 0.5;
    if (option(OPT_HOST_SPLIT, &ov) && ov > 0 && ov < 1) first_share = ov;
    double edge_w = 1.0 / 3.0;  // size of the first and the last chunk relative to the inner ones
    if (option(OPT_HOST_EDGE_WEIGHT, &ov) && ov > 0) edge_w = ov;
    HostCtx *ctx = host_ctx();
    if (!ctx) return cuda_fail(cudaGetLastError(), "mot3d_track_batch_host: stream/event/pinned scratch creation");

    // chunk-local graph_ptr arrays: pinned, so that their H2D copies are asynchronous like the rest
    if (ctx->gptr_cap < (size_t)(G + HOST_MAX_CHUNKS + 1)) {
        if (ctx->pinned_gptr) cudaFreeHost(ctx->pinned_gptr);
        ctx->pinned_gptr = nullptr;
        ctx->gptr_cap = 0;
        const size_t cap = (size_t)(G + HOST_MAX_CHUNKS + 1) * 2;
        M3_CUDA(cudaHostAlloc((void **)&ctx->pinned_gptr, cap * sizeof(int32_t), cudaHostAllocDefault));
        if (ctx->pinned_gbase) cudaFreeHost(ctx->pinned_gbase);
        ctx->pinned_gbase = nullptr;
        M3_CUDA(cudaHostAlloc((void **)&ctx->pinned_gbase, cap * sizeof(int64_t), cudaHostAllocDefault));
        ctx->gptr_cap = cap;
    }
    int32_t *h_gptr = ctx->pinned_gptr;
    int64_t *h_gbase = ctx->pinned_gbase;
    const bool tracing = option(OPT_HOST_TRACE, &ov) && ov != 0;
    if (tracing && !ctx->t0) {
        M3_CUDA(cudaEventCreate(&ctx->t0));
        for (int c = 0; c < HOST_MAX_CHUNKS; c++)
            for (int k = 0; k < 4; k++) M3_CUDA(cudaEventCreate(&ctx->trace[c][k]));
    }
    int32_t rc = MOT3D_OK;
    for (int chunks = want;; chunks = 1) {
        // cut at graph boundaries closest to equal detection counts
        int32_t cut[HOST_MAX_CHUNKS + 1];
        cut[0] = 0;
        int nc = 0;
        // a short first chunk gets the SMs going early, a short last one keeps the exposed D2H tail short
        double wsum = 0, wacc = 0, wgt[HOST_MAX_CHUNKS];
        for (int c = 0; c < chunks; c++) wsum += (wgt[c] = (chunks >= 3 && (c == 0 || c == chunks - 1)) ? edge_w : 1.0);
        if (chunks == 2) {  // the second chunk's solver and D2H are the exposed tail: it gets the smaller share
            wgt[0] = first_share;
            wgt[1] = 1.0 - first_share;
            wsum = 1.0;
        }
        for (int c = 1, g = 0; c <= chunks; c++) {
            wacc += wgt[c - 1];
            const int64_t target = (int64_t)((double)n * (wacc / wsum));
            while (g < G && hb->graph_ptr[g] < target) g++;
            if (c == chunks) g = G;
            if (g > cut[nc]) cut[++nc] = g;
        }
        Arena A(device_ws, ws_bytes);
        const bool pack_ptr = nc == 1 && !(option(OPT_HOST_NO_PACK, &ov) && ov != 0);
        int32_t *d_packed = nullptr;
        int64_t packed_total = 0;
        M3_CUDA(cudaEventRecord(ctx->start, user));
        M3_CUDA(cudaStreamWaitEvent(ctx->copy_in, ctx->start, 0));
        if (tracing) M3_CUDA(cudaEventRecord(ctx->t0, user));
        int64_t edges = 0;
        bool retry = false;
        rc = MOT3D_OK;
        int32_t *hg = h_gptr;
        for (int c = 0; c < nc && rc == MOT3D_OK; c++) {
            const int32_t ga = cut[c], gb = cut[c + 1];
            const int64_t d0 = hb->graph_ptr[ga], nd = hb->graph_ptr[gb] - d0;
            int32_t max_nodes = 0;
            int64_t key_base = 0;  // mot3d_make_keys: base = running sum of (frame span + max_jump + 1)
            int64_t *hb_base = h_gbase + ga;
            for (int g = ga; g <= gb; g++) {
                hg[g - ga] = (int32_t)(hb->graph_ptr[g] - d0);
                if (g < gb) {
                    const int32_t a = hb->graph_ptr[g], b = hb->graph_ptr[g + 1];
                    if (b - a > max_nodes) max_nodes = b - a;
                    hb_base[g - ga] = key_base;
                    const int64_t span = b > a ? (int64_t)hb->frame[b - 1] - (int64_t)hb->frame[a] : 0;
                    key_base += (span > 0 ? span : 0) + spec->max_jump + 1;
                }
            }
            if (key_base < 0 || key_base > (int64_t)0x7fffffff - MOT3D_MAX_JUMP - 1) {
                // (unsorted frames can make a span negative: the device check of the keys reports those)
                set_error("the frame spans of graphs [%d, %d) add up to %lld time keys, beyond 31 bits: split the batch",
                          ga, gb, (long long)key_base);
                rc = MOT3D_E_UNSUPPORTED;
                break;
            }
            // the arc buffer is shared out by detection count; a chunk that needs more sends the batch through again
            // in one piece, so the capacity the caller has to provide does not depend on the chunking
            const int64_t cap = nc == 1 ? edge_capacity : (int64_t)((double)edge_capacity * ((double)nd / (double)n));
            cudaStream_t st = nc == 1 ? user : ctx->s[c];
            if (nc > 1) M3_CUDA(cudaStreamWaitEvent(st, ctx->start, 0));
            int pieces = (int)(nd / HOST_PIECE_MIN_DETS);
            if (option(OPT_HOST_PIECES, &ov)) pieces = (int)ov;
            if (pieces > HOST_MAX_PIECES) pieces = HOST_MAX_PIECES;
            if (pieces < 1) pieces = 1;
            rc = enqueue_chunk(hb, spec, res, A, ga, gb, hg, hb_base, max_nodes, cap, ctx->pinned + 2 * c, st,
                               ctx->copy_in, ctx->in + c * HOST_MAX_PIECES, pieces, tracing ? ctx->trace[c] : nullptr,
                               ctx->done[0], pack_ptr ? &d_packed : nullptr);
            if (pack_ptr && rc == MOT3D_OK) {
                // one chunk: wait for the track counts only, then fetch sum(K + 1) packed entries into the tail of the
                // caller's track_ptr (behind the track_det copy that is already queued)
                M3_CUDA(cudaEventSynchronize(ctx->done[0]));
                for (int g = 0; g < G; g++) packed_total += (int64_t)res->track_count[g] + 1;
                if (packed_total > 0)
                    M3_CUDA(cudaMemcpyAsync(res->track_ptr + (n + G - packed_total), d_packed, (size_t)packed_total * 4,
                                            cudaMemcpyDeviceToHost, st));
            }
            if (nc > 1 && rc == MOT3D_OK) {
                M3_CUDA(cudaEventRecord(ctx->done[c], st));
                M3_CUDA(cudaStreamWaitEvent(user, ctx->done[c], 0));
            }
            hg += gb - ga + 1;
        }
        // everything enqueued so far has to drain before the workspace or h_gptr can be reused, error or not
        cudaError_t se = cudaSuccess;
        if (nc > 1)
            for (int c = 0; c < nc; c++) {
                cudaError_t e1 = cudaStreamSynchronize(ctx->s[c]);
                if (se == cudaSuccess) se = e1;
            }
        cudaError_t e2 = cudaStreamSynchronize(user);
        if (se == cudaSuccess) se = e2;
        if (rc == MOT3D_E_CAPACITY && !A.ok) {
            if (nc > 1) continue;  // the allowance per chunk did not cover this cut: one piece always fits
            set_error("device workspace too small: need %zu bytes", mot3d_track_batch_workspace_bytes(n, G, edge_capacity));
            break;
        }
        if (rc != MOT3D_OK) break;
        if (se != cudaSuccess) {
            rc = cuda_fail(se, "mot3d_track_batch_host");
            break;
        }
        if (pack_ptr && packed_total > 0) {
            // spread the packed entries out to graph_ptr[g] + g, ascending: the staging area at the tail never runs
            // ahead of the slice being written (sum_{h>=g} (K_h + 1) <= (n - graph_ptr[g]) + (G - g))
            const int32_t *src = res->track_ptr + (n + G - packed_total);
            for (int g = 0; g < G; g++) {
                const int32_t k1 = res->track_count[g] + 1;
                memmove(res->track_ptr + hb->graph_ptr[g] + g, src, (size_t)k1 * 4);
                src += k1;
            }
        }
        if (tracing)
            for (int c = 0; c < nc; c++) {
                float t[4] = {0, 0, 0, 0};
                for (int k = 0; k < 4; k++) cudaEventElapsedTime(&t[k], ctx->t0, ctx->trace[c][k]);
                fprintf(stderr, "mot3d host pipeline: chunk %d/%d graphs [%d, %d): H2D done %.3f ms, build done %.3f, "
                                "solve+tracks done %.3f, D2H done %.3f\n", c, nc, cut[c], cut[c + 1], t[0], t[1], t[2], t[3]);
            }
        for (int c = 0; c < nc; c++) {
            const int64_t nd = (int64_t)hb->graph_ptr[cut[c + 1]] - hb->graph_ptr[cut[c]];
            const int64_t cap = nc == 1 ? edge_capacity : (int64_t)((double)edge_capacity * ((double)nd / (double)n));
            edges += ctx->pinned[2 * c + 1];
            const int32_t flag = ctx->pinned[2 * c];
            if (flag & MOT3D_ERR_UNSORTED) {
                set_error("the frames of a graph are not non-decreasing (graphs [%d, %d))", cut[c], cut[c + 1]);
                rc = MOT3D_E_INVALID;
            } else if (flag & (MOT3D_ERR_KEY_RANGE | MOT3D_ERR_WINDOW)) {
                set_error("time keys beyond 31 bits or more than 2^24 detections inside one time window");
                rc = MOT3D_E_UNSUPPORTED;
            }
            if ((flag & MOT3D_ERR_ARC_CAPACITY) || ctx->pinned[2 * c + 1] > cap) retry = true;
        }
        if (rc != MOT3D_OK) break;
        res->n_edges = edges;
        // what travelled: graph_ptr + key bases, frame, positions, confidences [, flags, boxes] in; counts, totals,
        // track_det, the (packed) track_ptr and two flags per chunk out
        res->h2d_bytes += (int64_t)(G + nc) * 4 + (int64_t)G * 8 + n * (4 + 8 * hb->dim + 8) + (hb->flags ? n : 0) +
                          (spec->use_bbox ? n * 32 : 0);
        res->d2h_bytes += (int64_t)G * 12 + n * 4 + (pack_ptr ? packed_total * 4 : (n + G) * 4) + (int64_t)nc * 8;
        if (retry) {
            if (nc > 1 && edges <= edge_capacity) continue;  // the whole batch fits: run it in one piece
            set_error("edge capacity %lld too small: the batch has %lld transition arcs", (long long)edge_capacity,
                      (long long)edges);
            rc = MOT3D_E_CAPACITY;
        }
        break;
    }
    return rc;
}

}  // extern "C"
