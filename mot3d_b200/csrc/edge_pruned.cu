// Spatially pruned variant of the windowed all-pairs kernel (same outputs as k_edge_build, csrc/edge_build.cu).
//
// At the reference's densities only ~1 % of the candidate pairs inside the time window survive the distance test
// (weight_distance_detections_2d returns None for dist > max_distance, mot3d/weight_functions.py:21-23), so testing
// every pair makes the kernel FP64-issue bound at a few per cent of the HBM roofline. Here every frame gets an
// auxiliary copy of its detections sorted by x (k_frame_sort: rank by counting, stable); a row then binary-searches
// the x-window [x - d, x + d] in each of the <= max_jump frames it may link to and tests only those candidates
// exactly (same FP64 expression and threshold as the brute-force kernel, so the arc set is identical). Arcs of one
// (row, frame) come out in x order and are put back in detection order (what the reference's loop emits,
// mot3d/mot3d.py:111-127) by an insertion sort over the 1-3 entries involved.
#include "common.cuh"

namespace m3 {

constexpr int EP_TPB = 128;
constexpr int EP_CAP = 1024;   // candidates staged per tile; larger windows read the global arrays directly
constexpr int EP_TABLE = 96;   // distinct time keys a tile may see

struct PrunedParams {
    int64_t n;
    int32_t dim;
    int32_t max_jump;
    const int32_t *tkey;
    const double *pos;        // [dim][n] original order (rows)
    const double *spos;       // [dim][n] x-sorted inside every frame (candidates)
    const int32_t *sperm;     // [n] original index of every sorted entry
    double sq_dist_max;
    double x_window;          // >= sqrt(sq_dist_max)
    int32_t *row_count;
    const int32_t *row_ptr;
    int64_t edge_capacity;
    int32_t *col;
    int64_t *cost;
    int32_t *overflow;
};

template <bool UPPER>
__device__ __forceinline__ int64_t bound_i32(const int32_t *a, int64_t lo, int64_t hi, int64_t key) {
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        const int64_t v = a[mid];
        if (UPPER ? (v <= key) : (v < key))
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}

// rank of every detection inside its frame by (x, index): sorted copies of the positions + permutation
__global__ void __launch_bounds__(128) k_frame_sort(int64_t n, int dim, const int32_t *__restrict__ tkey,
                                                    const double *__restrict__ pos, double *__restrict__ spos,
                                                    int32_t *__restrict__ sperm) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int k = tkey[i];
    const int64_t lo = bound_i32<false>(tkey, 0, i + 1, k);
    const int64_t hi = bound_i32<true>(tkey, i, n, k);
    const double x = pos[i];
    int rank = 0;
    for (int64_t j = lo; j < hi; j++) {
        const double xj = pos[j];
        rank += (xj < x || (xj == x && j < i)) ? 1 : 0;
    }
    const int64_t o = lo + rank;
    sperm[o] = (int32_t)i;
    for (int d = 0; d < dim; d++) spos[(size_t)d * n + o] = pos[(size_t)d * n + i];
}

template <int DIR, bool FILL>
__global__ void __launch_bounds__(EP_TPB) k_edge_build_pruned(const PrunedParams p) {
    extern __shared__ __align__(16) unsigned char ep_smem[];
    __shared__ int64_t s_range[2];
    __shared__ int s_seg_lo[EP_TABLE], s_seg_hi[EP_TABLE];
    __shared__ __align__(8) uint64_t s_bar;

    const int64_t n = p.n;
    const int64_t r0 = (int64_t)blockIdx.x * EP_TPB;
    const int64_t r1 = (r0 + EP_TPB < n) ? r0 + EP_TPB : n;
    const int64_t r = r0 + threadIdx.x;
    const bool active = r < n;
    const int J = p.max_jump;
    const bool has_z = p.dim == 3;

    // candidate index range of the tile (the sorted copies permute detections inside frames only)
    const int64_t k_first = p.tkey[r0], k_last = p.tkey[r1 - 1];
    const int64_t kmin = (DIR == MOT3D_DIR_OUT) ? k_first + 1 : k_first - J;
    const int64_t kmax = (DIR == MOT3D_DIR_OUT) ? k_last + J : k_last - 1;
    if (threadIdx.x == 0) {
        s_range[0] = bound_i32<false>(p.tkey, 0, n, kmin);
        s_range[1] = bound_i32<true>(p.tkey, 0, n, kmax);
        mbar_init(&s_bar, 1);
        fence_mbar_init();
    }
    for (int t = threadIdx.x; t < EP_TABLE; t += EP_TPB) {
        s_seg_lo[t] = 0;
        s_seg_hi[t] = 0;
    }
    __syncthreads();
    const int64_t lo = s_range[0], hi = s_range[1];
    const int m = (int)(hi - lo);
    const bool staged = m <= EP_CAP;
    const bool tabled = (kmax - kmin) < EP_TABLE;
    // staging buffers: every array keeps 4 spare slots so that the 16-byte aligned body of its slice can go through
    // the TMA unit (cp.async.bulk, SASS UBLKCP) with the element alignment of the global address preserved
    double *b_x = (double *)ep_smem;
    double *b_y = b_x + EP_CAP + 4;
    double *b_z = b_y + EP_CAP + 4;  // only touched when has_z (the launch sizes the buffer accordingly)
    int32_t *b_perm = (int32_t *)(b_y + EP_CAP + 4 + (has_z ? EP_CAP + 4 : 0));
    int32_t *b_key = b_perm + EP_CAP + 4;
    const double *s_x = b_x, *s_y = b_y, *s_z = b_z;
    const int32_t *s_perm = b_perm, *s_key = b_key;
    if (staged && m > 0) {
        uint32_t bulk = 0;
        s_x = b_x + stage_slice(p.spos, b_x, lo, m, &s_bar, &bulk);
        s_y = b_y + stage_slice(p.spos + n, b_y, lo, m, &s_bar, &bulk);
        if (has_z) s_z = b_z + stage_slice(p.spos + 2 * n, b_z, lo, m, &s_bar, &bulk);
        s_perm = b_perm + stage_slice(p.sperm, b_perm, lo, m, &s_bar, &bulk);
        s_key = b_key + stage_slice(p.tkey, b_key, lo, m, &s_bar, &bulk);
        if (threadIdx.x == 0) mbar_expect_tx(&s_bar, bulk);
        __syncthreads();  // ragged head/tail stores visible
        mbar_wait(&s_bar, 0);
    }
    __syncthreads();
    // frame segments of the candidate range, indexed by key - kmin
    if (tabled) {
        for (int c = threadIdx.x; c < m; c += EP_TPB) {
            const int k = staged ? s_key[c] : p.tkey[lo + c];
            const int kp = c > 0 ? (staged ? s_key[c - 1] : p.tkey[lo + c - 1]) : k - 1;
            const int kn = c + 1 < m ? (staged ? s_key[c + 1] : p.tkey[lo + c + 1]) : k + 1;
            if (kp != k) s_seg_lo[k - kmin] = c;
            if (kn != k) s_seg_hi[k - kmin] = c + 1;
        }
    }
    __syncthreads();
    if (!active) return;

    const int key_r = p.tkey[r];
    const double xr = p.pos[r], yr = p.pos[n + r];
    const double zr = has_z ? p.pos[2 * n + r] : 0.0;
    const double xlo = xr - p.x_window, xhi = xr + p.x_window;
    const double *gx = p.spos + lo, *gy = p.spos + n + lo, *gz = p.spos + 2 * n + lo;
    const int32_t *gperm = p.sperm + lo, *gkey = p.tkey + lo;
    int cnt = 0;
    const int64_t out_base = FILL ? p.row_ptr[r] : 0;

    for (int jj = 1; jj <= J; jj++) {
        // frames in ascending time order, so that the arcs of a row ascend by detection index
        const int kt = (DIR == MOT3D_DIR_OUT) ? key_r + jj : key_r - (J + 1 - jj);
        int a, b;
        if (tabled) {
            const int t = kt - (int)kmin;
            a = s_seg_lo[t];
            b = s_seg_hi[t];
        } else {
            a = (int)bound_i32<false>(gkey, 0, m, kt);
            b = (int)bound_i32<true>(gkey, a, m, kt);
        }
        if (b <= a) continue;
        // first candidate with x >= xr - window
        int l = a, h = b;
        while (l < h) {
            const int mid = (l + h) >> 1;
            const double v = staged ? s_x[mid] : gx[mid];
            if (v < xlo)
                l = mid + 1;
            else
                h = mid;
        }
        const int first = cnt;
        for (int c = l; c < b; c++) {
            const double xc = staged ? s_x[c] : gx[c];
            if (xc > xhi) break;
            const double dx = xr - xc;
            const double dy = yr - (staged ? s_y[c] : gy[c]);
            double s = dx * dx;  // sum([dx^2, dy^2, dz^2]) left to right, as the reference's euclidean()
            s = s + dy * dy;
            if (has_z) {
                const double dz = zr - (staged ? s_z[c] : gz[c]);
                s = s + dz * dz;
            }
            if (s <= p.sq_dist_max) {
                if (FILL) {
                    const int other = staged ? s_perm[c] : gperm[c];
                    int64_t e = out_base + cnt;
                    if (e < p.edge_capacity) {
                        // insertion into the (short) run of this frame's arcs, ascending detection index
                        const int64_t e_first = out_base + first;
                        while (e > e_first && p.col[e - 1] > other) {
                            p.col[e] = p.col[e - 1];
                            p.cost[e] = p.cost[e - 1];
                            e--;
                        }
                        p.col[e] = other;
                        p.cost[e] = __double_as_longlong(s);
                    } else if (p.overflow) {
                        *p.overflow = 1;
                    }
                }
                cnt++;
            }
        }
    }
    if (!FILL) p.row_count[r] = cnt;
}

}  // namespace m3

using namespace m3;

// internal entry points used by mot3d_edge_count / mot3d_edge_fill (edge_build.cu) when sorted copies are supplied
namespace m3 {

int32_t launch_frame_sort(const mot3d_dets *d, cudaStream_t st) {
    k_frame_sort<<<(unsigned)((d->n + 127) / 128), 128, 0, st>>>(d->n, d->dim, d->tkey, d->pos, d->sorted_pos,
                                                                 d->sorted_perm);
    M3_LAUNCH_CHECK("k_frame_sort");
    return MOT3D_OK;
}

int32_t launch_pruned(const mot3d_dets *d, const mot3d_cost_spec *s, int direction, bool fill, int32_t *row_count,
                      const int32_t *row_ptr, int64_t edge_capacity, int32_t *col, int64_t *cost, int32_t *overflow,
                      cudaStream_t st) {
    PrunedParams p;
    p.n = d->n;
    p.dim = d->dim;
    p.max_jump = s->max_jump;
    p.tkey = d->tkey;
    p.pos = d->pos;
    p.spos = d->sorted_pos;
    p.sperm = d->sorted_perm;
    p.sq_dist_max = s->sq_dist_max;
    // a superset of the x-interval that can hold survivors: dx^2 <= s <= sq_dist_max
    const double w = s->sq_dist_max > 0 ? sqrt(s->sq_dist_max) : 0.0;
    p.x_window = w * (1.0 + 1e-12) + 1e-300;
    p.row_count = row_count;
    p.row_ptr = row_ptr;
    p.edge_capacity = edge_capacity;
    p.col = col;
    p.cost = cost;
    p.overflow = overflow;
    const size_t smem = (size_t)(EP_CAP + 4) * (8 + 8 + (d->dim == 3 ? 8 : 0) + 4 + 4) + 16;
    const unsigned grid = (unsigned)((d->n + EP_TPB - 1) / EP_TPB);
#define M3_EP_LAUNCH(DIRV, FILLV)                                                                                  \
    do {                                                                                                           \
        M3_CUDA(cudaFuncSetAttribute(k_edge_build_pruned<DIRV, FILLV>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                     (int)smem));                                                                  \
        k_edge_build_pruned<DIRV, FILLV><<<grid, EP_TPB, smem, st>>>(p);                                           \
    } while (0)
    if (direction == MOT3D_DIR_IN) {
        if (fill)
            M3_EP_LAUNCH(MOT3D_DIR_IN, true);
        else
            M3_EP_LAUNCH(MOT3D_DIR_IN, false);
    } else {
        if (fill)
            M3_EP_LAUNCH(MOT3D_DIR_OUT, true);
        else
            M3_EP_LAUNCH(MOT3D_DIR_OUT, false);
    }
#undef M3_EP_LAUNCH
    M3_LAUNCH_CHECK("k_edge_build_pruned");
    return MOT3D_OK;
}

}  // namespace m3
