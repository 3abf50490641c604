// C-ABI plumbing: version, thread-local error text, and the one-call host-buffer entry point
// (pack -> H2D -> keys -> count -> scan -> fill -> solve -> tracks -> D2H).
#include <stdarg.h>
#include <string.h>

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

size_t mot3d_track_batch_workspace_bytes(int64_t n, int32_t n_graphs, int64_t edge_capacity) {
    if (n < 0 || n_graphs < 0 || edge_capacity < 0) return 0;
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
    b += align256((size_t)n1 * 4);              // row_count
    b += align256((size_t)(n1 + 1) * 4);        // row_ptr
    b += align256(mot3d_scan_workspace_bytes(n1));
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
    return b + 4096;
}

int32_t mot3d_track_batch_host(const mot3d_host_batch *hb, const mot3d_cost_spec *spec, mot3d_host_result *res,
                               void *device_ws, size_t ws_bytes, int64_t edge_capacity, void *stream) {
    M3_CHECK_ARG(hb && spec && res, "NULL batch/spec/result");
    M3_CHECK_ARG(hb->n_graphs >= 0 && hb->n >= 0, "negative sizes");
    M3_CHECK_ARG(!spec->use_hist, "mot3d_track_batch_host: colour histograms go through the device entry points");
    res->n_edges = 0;
    if (hb->n == 0 || hb->n_graphs == 0) return MOT3D_OK;
    M3_CHECK_ARG(hb->graph_ptr && hb->frame && hb->pos && hb->conf, "NULL host input");
    M3_CHECK_ARG(!spec->use_bbox || hb->bbox, "use_bbox set but bbox is NULL");
    M3_CHECK_ARG(res->track_count && res->track_ptr && res->track_det && res->total_cost, "NULL host output");
    M3_CHECK_ARG(device_ws, "device_ws is NULL");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n = hb->n;
    const int32_t G = hb->n_graphs;
    int32_t max_nodes = 0;
    for (int g = 0; g < G; g++) {
        int32_t c = hb->graph_ptr[g + 1] - hb->graph_ptr[g];
        M3_CHECK_ARG(c >= 0, "graph_ptr must be non-decreasing");
        if (c > max_nodes) max_nodes = c;
    }
    M3_CHECK_ARG(hb->graph_ptr[0] == 0 && hb->graph_ptr[G] == n, "graph_ptr must span [0, n]");

    Arena A(device_ws, ws_bytes);
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
    int32_t *d_rowcnt = A.take<int32_t>(n);
    int32_t *d_rowptr = A.take<int32_t>(n + 1);
    size_t scan_b = mot3d_scan_workspace_bytes(n);
    char *d_scan = A.take<char>((int64_t)scan_b);
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
    if (!A.ok) {
        set_error("device workspace too small: need %zu bytes", mot3d_track_batch_workspace_bytes(n, G, edge_capacity));
        return MOT3D_E_CAPACITY;
    }

    M3_CUDA(cudaMemcpyAsync(d_gptr, hb->graph_ptr, (size_t)(G + 1) * 4, cudaMemcpyHostToDevice, st));
    M3_CUDA(cudaMemcpyAsync(d_frame, hb->frame, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    M3_CUDA(cudaMemcpyAsync(d_pos, hb->pos, (size_t)n * 8 * hb->dim, cudaMemcpyHostToDevice, st));
    M3_CUDA(cudaMemcpyAsync(d_conf, hb->conf, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    if (d_flags) M3_CUDA(cudaMemcpyAsync(d_flags, hb->flags, (size_t)n, cudaMemcpyHostToDevice, st));
    if (d_bbox) M3_CUDA(cudaMemcpyAsync(d_bbox, hb->bbox, (size_t)n * 32, cudaMemcpyHostToDevice, st));
    M3_CUDA(cudaMemsetAsync(d_over, 0, 4, st));

    int32_t rc;
    if ((rc = mot3d_make_keys(d_frame, d_gptr, G, spec->max_jump, d_tkey, d_gbase, st))) return rc;
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
    if ((rc = mot3d_node_costs(&d, spec, d_en, d_cn, d_ex, st))) return rc;
    if ((rc = mot3d_frame_sort(&d, st))) return rc;
    if ((rc = mot3d_edge_count(&d, spec, MOT3D_DIR_IN, d_rowcnt, st))) return rc;
    if ((rc = mot3d_exclusive_scan_i32(d_rowcnt, n, d_rowptr, d_scan, scan_b, st))) return rc;
    if ((rc = mot3d_edge_fill(&d, spec, MOT3D_DIR_IN, d_rowptr, edge_capacity, d_col, d_cost, nullptr, d_over, st)))
        return rc;
    if ((rc = mot3d_solve_batch(G, n, d_gptr, max_nodes, d_tkey, edge_capacity, d_rowptr, d_col, d_cost, d_en, d_cn, d_ex, d_next, d_prev,
                                d_np, d_total, nullptr, nullptr, d_solve, solve_b, st)))
        return rc;
    if ((rc = mot3d_extract_tracks(G, d_gptr, d_next, d_prev, d_tcount, d_tptr, d_tdet, st))) return rc;

    int32_t h_over = 0, h_edges = 0;
    M3_CUDA(cudaMemcpyAsync(res->track_count, d_tcount, (size_t)G * 4, cudaMemcpyDeviceToHost, st));
    M3_CUDA(cudaMemcpyAsync(res->track_ptr, d_tptr, (size_t)(n + G) * 4, cudaMemcpyDeviceToHost, st));
    M3_CUDA(cudaMemcpyAsync(res->track_det, d_tdet, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    M3_CUDA(cudaMemcpyAsync(res->total_cost, d_total, (size_t)G * 8, cudaMemcpyDeviceToHost, st));
    M3_CUDA(cudaMemcpyAsync(&h_over, d_over, 4, cudaMemcpyDeviceToHost, st));
    M3_CUDA(cudaMemcpyAsync(&h_edges, d_rowptr + n, 4, cudaMemcpyDeviceToHost, st));
    M3_CUDA(cudaStreamSynchronize(st));
    res->n_edges = h_edges;
    if (h_over || h_edges > edge_capacity) {
        set_error("edge capacity %lld too small: the batch has %d transition arcs", (long long)edge_capacity, h_edges);
        return MOT3D_E_CAPACITY;
    }
    return MOT3D_OK;
}

}  // extern "C"
