// Tracklet-linking graph build (second pass of the reference pipelines): arcs between tracklets.
//
// Reference semantics (cvlab-epfl/mot3d):
//   build_graph on DetectionTracklet lists   mot3d/mot3d.py:106-141  no sort; every ordered pair (t1, t2) with
//                                            0 < t2.tail.index - t1.head.index <= max_jump (mot3d/types.py:67-68)
//   weight_distance_tracklets_2d / _3d       mot3d/weight_functions.py:80-123, 128-160
//   linear_motion                            mot3d/distance_functions.py:47-63: degree-1 least-squares fit through one
//                                            end, evaluated at the other end's frames, mean L2 error, both directions
//   color_histogram_similarity               mot3d/distance_functions.py:21-34 (median histograms of the two ends)
// The per-end line fits are computed once per tracklet (k_tracklet_fit, closed form) instead of once per pair.
// One warp per row (tail tracklet t1), lanes stride over t2, warp-ballot compaction keeps the reference's arc order.
#include "common.cuh"

namespace m3 {

struct TrkParams {
    int32_t n;
    int32_t dim;
    int32_t max_jump;
    int32_t use_hist;
    int32_t check_access;  // weight_distance_tracklets_2d refuses arcs into t2.tail.access_point (:101-102)
    int32_t C, B;
    int32_t pad;
    const int32_t *head_ptr, *tail_ptr;  // [n+1] windows of the two ends
    const int32_t *head_idx, *tail_idx;  // frame indexes of the window points (ascending)
    const double *head_pos, *tail_pos;   // [dim][H] / [dim][Tl] planar
    int64_t n_head, n_tail;
    double *head_fit, *tail_fit;         // [n][dim][2] slope, intercept
    const uint8_t *tail_access;          // [n] or NULL
    const uint8_t *head_has_hist, *tail_has_hist;  // [n] or NULL (= every end has one)
    const double *head_hist, *tail_hist;           // [n][C][B]
    const int32_t *view_ptr, *view_id;             // per-view entries of 3-D tracklets (NULL = single histogram)
    const double *view_head_hist, *view_tail_hist; // [entries][C][B]
    double sq_dist_max;  // < 0: no max_distance gate
    double sigma_motion_sq, sigma_hist_sq, alpha, cutoff_motion, cutoff_appearance;
    int32_t *row_count;
    const int32_t *row_ptr;
    int64_t edge_capacity;
    int32_t *col;
    int64_t *cost;
    double *weight;
    int32_t *overflow;
};

// least squares y = slope * x + intercept over the window (np.polyfit(idx, pos, 1), deg 1)
__global__ void k_tracklet_fit(TrkParams p) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * p.n * p.dim) return;
    const int d = t % p.dim;
    const int i = (t / p.dim) % p.n;
    const bool tail = t >= p.n * p.dim;
    const int32_t *ptr = tail ? p.tail_ptr : p.head_ptr;
    const int32_t *idx = tail ? p.tail_idx : p.head_idx;
    const double *pos = (tail ? p.tail_pos : p.head_pos) + (size_t)d * (tail ? p.n_tail : p.n_head);
    double *fit = (tail ? p.tail_fit : p.head_fit) + ((size_t)i * p.dim + d) * 2;
    const int a = ptr[i], b = ptr[i + 1];
    const int m = b - a;
    double slope = 0.0, icpt = 0.0;
    if (m == 1) {
        // rank-deficient: numpy's lstsq returns the minimum-norm solution of the column-scaled system
        const double x = (double)idx[a], y = pos[a];
        if (x != 0.0) {
            slope = y / (2.0 * x);
            icpt = y / 2.0;
        } else {
            icpt = y;
        }
    } else if (m > 1) {
        double sx = 0, sy = 0;
        for (int k = a; k < b; k++) {
            sx += (double)idx[k];
            sy += pos[k];
        }
        const double mx = sx / m, my = sy / m;
        double sxx = 0, sxy = 0;
        for (int k = a; k < b; k++) {
            const double dx = (double)idx[k] - mx;
            sxx += dx * dx;
            sxy += dx * (pos[k] - my);
        }
        if (sxx > 0) {
            slope = sxy / sxx;
            icpt = my - slope * mx;
        } else {
            icpt = my;
        }
    }
    fit[0] = slope;
    fit[1] = icpt;
}

// mean over the points of `b_end` of || line_a(idx) - pos ||
__device__ __forceinline__ double one_way_deviation(const TrkParams &p, const double *fit_a, const int32_t *idx,
                                                    const double *pos, int64_t plane, int a, int b) {
    double acc = 0.0;
    for (int k = a; k < b; k++) {
        const double x = (double)idx[k];
        double s = 0.0;
        for (int d = 0; d < p.dim; d++) {
            const double pred = fit_a[2 * d] * x + fit_a[2 * d + 1];
            const double e = pred - pos[(size_t)d * plane + k];
            s += e * e;
        }
        acc += sqrt(s);
    }
    return acc / (double)(b - a);
}

// weight of the arc t1 -> t2, or false when the reference returns None / the pair is outside the time window
// 1 - color_histogram_similarity(h1, h2) (mot3d/distance_functions.py:21-34, unit channel weights)
__device__ __forceinline__ double hist_dissimilarity(const TrkParams &p, const double *h1, const double *h2) {
    double scs = 0.0;
    for (int c = 0; c < p.C; c++) {
        double ab = 0, aa = 0, bb = 0;
        for (int b = 0; b < p.B; b++) {
            const double u = h1[c * p.B + b], v = h2[c * p.B + b];
            ab += u * v;
            aa += u * u;
            bb += v * v;
        }
        scs += ab / sqrt(aa * bb);
    }
    return 1.0 - scs / (double)p.C;
}

__device__ __forceinline__ bool tracklet_weight(const TrkParams &p, int i, int j, double *w_out) {
    const int hb = p.head_ptr[i + 1], ha = p.head_ptr[i];
    const int ta = p.tail_ptr[j], tb = p.tail_ptr[j + 1];
    if (hb <= ha || tb <= ta) return false;
    const int jump = p.tail_idx[ta] - p.head_idx[hb - 1];  // t2.tail.index - t1.head.index
    if (jump <= 0 || jump > p.max_jump) return false;
    if (p.sq_dist_max >= 0.0) {
        double s = 0.0;
        for (int d = 0; d < p.dim; d++) {
            const double e = p.head_pos[(size_t)d * p.n_head + hb - 1] - p.tail_pos[(size_t)d * p.n_tail + ta];
            s += e * e;
        }
        if (s > p.sq_dist_max) return false;
    }
    if (p.check_access && p.tail_access && p.tail_access[j]) return false;
    const double fwd = one_way_deviation(p, p.head_fit + (size_t)i * p.dim * 2, p.tail_idx, p.tail_pos, p.n_tail, ta, tb);
    const double bwd = one_way_deviation(p, p.tail_fit + (size_t)j * p.dim * 2, p.head_idx, p.head_pos, p.n_head, ha, hb);
    const double dev = fwd + bwd;
    const double wm = exp(-(dev * dev) / p.sigma_motion_sq);
    if (wm < p.cutoff_motion) return false;
    if (p.use_hist && p.view_ptr) {
        // 3-D tracklets: mean over the views both have (t1's dict order) of 1 - NCC(head of t1's 2-D tracklet, tail of
        // t2's) (mot3d/weight_functions.py:146-158); np.mean of < 8 values adds them left to right
        double sum = 0.0;
        int shared = 0;
        for (int u = p.view_ptr[i]; u < p.view_ptr[i + 1]; u++) {
            const int id = p.view_id[u];
            for (int v = p.view_ptr[j]; v < p.view_ptr[j + 1]; v++)
                if (p.view_id[v] == id) {
                    sum = sum + hist_dissimilarity(p, p.view_head_hist + (size_t)u * p.C * p.B,
                                                   p.view_tail_hist + (size_t)v * p.C * p.B);
                    shared++;
                    break;
                }
        }
        if (shared == 0) {  // nan in the reference; reported through the overflow flag (bit 1) by the fill pass
            *w_out = __longlong_as_double(0x7ff8000000000000ll);
            return true;
        }
        const double mc = sum / (double)shared;
        const double wa = exp(-(mc * mc) / p.sigma_hist_sq);
        if (wa < p.cutoff_appearance) return false;
        *w_out = -((1.0 - p.alpha) * wm + p.alpha * wa);
        return true;
    }
    const bool hist = p.use_hist && (!p.head_has_hist || p.head_has_hist[i]) && (!p.tail_has_hist || p.tail_has_hist[j]);
    if (!hist) {
        *w_out = -wm;
        return true;
    }
    const double chs = hist_dissimilarity(p, p.head_hist + (size_t)i * p.C * p.B, p.tail_hist + (size_t)j * p.C * p.B);
    const double wa = exp(-(chs * chs) / p.sigma_hist_sq);
    if (wa < p.cutoff_appearance) return false;
    *w_out = -((1.0 - p.alpha) * wm + p.alpha * wa);
    return true;
}

template <bool FILL>
__global__ void __launch_bounds__(128) k_tracklet_edges(TrkParams p) {
    const int i = blockIdx.x * (blockDim.x >> 5) + warp_id();
    if (i >= p.n) return;
    const int lane = lane_id();
    int cnt = 0;
    const int64_t base = FILL ? p.row_ptr[i] : 0;
    for (int j0 = 0; j0 < p.n; j0 += 32) {
        const int j = j0 + lane;
        double w = 0.0;
        const bool keep = j < p.n && j != i && tracklet_weight(p, i, j, &w);
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (FILL && keep) {
            const int64_t e = base + cnt + __popc(m & ((1u << lane) - 1u));
            if (e < p.edge_capacity) {
                const bool nan = w != w;
                p.col[e] = j;
                p.cost[e] = nan ? 0 : quantise_1e7(w);
                if (p.weight) p.weight[e] = w;
                if (nan && p.overflow) atomicOr(p.overflow, 2);
            } else if (p.overflow) {
                atomicOr(p.overflow, 1);
            }
        }
        cnt += __popc(m);
    }
    if (!FILL && lane == 0) p.row_count[i] = cnt;
}

static int32_t fill_trk(TrkParams &p, const mot3d_tracklets *t, const mot3d_tracklet_spec *s) {
    M3_CHECK_ARG(t && s, "tracklets/spec is NULL");
    M3_CHECK_ARG(t->n >= 0 && (t->dim == 2 || t->dim == 3), "bad tracklet batch");
    M3_CHECK_ARG(s->max_jump >= 1, "max_jump must be >= 1");
    M3_CHECK_ARG(t->n == 0 || (t->head_ptr && t->tail_ptr && t->head_idx && t->tail_idx && t->head_pos && t->tail_pos &&
                               t->head_fit_ws && t->tail_fit_ws),
                 "NULL tracklet array");
    M3_CHECK_ARG(!s->use_hist || t->view_ptr ||
                     (t->head_hist && t->tail_hist && t->hist_channels > 0 && t->hist_bins > 0),
                 "use_hist set but histograms are NULL/empty");
    M3_CHECK_ARG(!(s->use_hist && t->view_ptr) ||
                     (t->view_id && t->view_head_hist && t->view_tail_hist && t->hist_channels > 0 && t->hist_bins > 0),
                 "view mode: view_id / view_head_hist / view_tail_hist is NULL or the histogram shape is empty");
    p.n = t->n;
    p.dim = t->dim;
    p.max_jump = s->max_jump;
    p.use_hist = s->use_hist;
    p.check_access = s->check_access_point;
    p.C = t->hist_channels;
    p.B = t->hist_bins;
    p.pad = 0;
    p.head_ptr = t->head_ptr;
    p.tail_ptr = t->tail_ptr;
    p.head_idx = t->head_idx;
    p.tail_idx = t->tail_idx;
    p.head_pos = t->head_pos;
    p.tail_pos = t->tail_pos;
    p.n_head = t->n_head_points;
    p.n_tail = t->n_tail_points;
    p.head_fit = t->head_fit_ws;
    p.tail_fit = t->tail_fit_ws;
    p.tail_access = t->tail_access_point;
    p.head_has_hist = t->head_has_hist;
    p.tail_has_hist = t->tail_has_hist;
    p.head_hist = t->head_hist;
    p.tail_hist = t->tail_hist;
    p.view_ptr = t->view_ptr;
    p.view_id = t->view_id;
    p.view_head_hist = t->view_head_hist;
    p.view_tail_hist = t->view_tail_hist;
    p.sq_dist_max = s->sq_dist_max;
    p.sigma_motion_sq = s->sigma_motion_sq;
    p.sigma_hist_sq = s->sigma_hist_sq;
    p.alpha = s->alpha;
    p.cutoff_motion = s->cutoff_motion;
    p.cutoff_appearance = s->cutoff_appearance;
    p.row_count = nullptr;
    p.row_ptr = nullptr;
    p.edge_capacity = 0;
    p.col = nullptr;
    p.cost = nullptr;
    p.weight = nullptr;
    p.overflow = nullptr;
    return MOT3D_OK;
}

}  // namespace m3

using namespace m3;

extern "C" {

int32_t mot3d_tracklet_edge_count(const mot3d_tracklets *trk, const mot3d_tracklet_spec *spec, int32_t *row_count_out,
                                  void *stream) {
    TrkParams p;
    int32_t rc = fill_trk(p, trk, spec);
    if (rc) return rc;
    if (p.n == 0) return MOT3D_OK;
    M3_CHECK_ARG(row_count_out, "row_count_out is NULL");
    p.row_count = row_count_out;
    cudaStream_t st = (cudaStream_t)stream;
    k_tracklet_fit<<<(2 * p.n * p.dim + 127) / 128, 128, 0, st>>>(p);
    M3_LAUNCH_CHECK("k_tracklet_fit");
    k_tracklet_edges<false><<<(p.n + 3) / 4, 128, 0, st>>>(p);
    M3_LAUNCH_CHECK("k_tracklet_edges<count>");
    return MOT3D_OK;
}

int32_t mot3d_tracklet_edge_fill(const mot3d_tracklets *trk, const mot3d_tracklet_spec *spec, const int32_t *row_ptr,
                                 int64_t edge_capacity, int32_t *col_out, int64_t *cost_out, double *weight_out,
                                 int32_t *overflow_flag, void *stream) {
    TrkParams p;
    int32_t rc = fill_trk(p, trk, spec);
    if (rc) return rc;
    if (p.n == 0) return MOT3D_OK;
    M3_CHECK_ARG(row_ptr && edge_capacity >= 0, "row_ptr is NULL / negative capacity");
    M3_CHECK_ARG(edge_capacity == 0 || (col_out && cost_out), "col_out/cost_out is NULL");
    p.row_ptr = row_ptr;
    p.edge_capacity = edge_capacity;
    p.col = col_out;
    p.cost = cost_out;
    p.weight = weight_out;
    p.overflow = overflow_flag;
    // the fits written by mot3d_tracklet_edge_count are still in the workspace (same stream order)
    k_tracklet_edges<true><<<(p.n + 3) / 4, 128, 0, (cudaStream_t)stream>>>(p);
    M3_LAUNCH_CHECK("k_tracklet_edges<fill>");
    return MOT3D_OK;
}

}  // extern "C"
