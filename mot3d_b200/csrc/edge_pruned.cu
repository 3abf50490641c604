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

// The same bound found by one warp: 32 probes per step shrink the range 33-fold, so a batch of millions of keys takes
// 5 dependent loads instead of 23. All 32 lanes must call; every lane returns the result.
template <bool UPPER>
__device__ __forceinline__ int64_t warp_bound_i32(const int32_t *a, int64_t lo, int64_t hi, int64_t key) {
    const int lane = threadIdx.x & 31;
    while (hi - lo > 32) {
        const int64_t span = hi - lo;
        const int64_t pos = lo + (span * (lane + 1)) / 33;  // lo < pos < hi, non-decreasing in lane
        const int64_t v = a[pos];
        const bool below = UPPER ? (v <= key) : (v < key);  // the bound lies beyond pos
        const unsigned m = __ballot_sync(0xffffffffu, below);
        const int c = __popc(m);  // probes 0 .. c-1 are below (the keys are sorted)
        const int64_t new_lo = c > 0 ? __shfl_sync(0xffffffffu, pos, c - 1) + 1 : lo;
        const int64_t new_hi = c < 32 ? __shfl_sync(0xffffffffu, pos, c < 32 ? c : 31) : hi;
        lo = new_lo;
        hi = new_hi;
    }
    const int64_t pos = lo + lane;
    const bool below = pos < hi && (UPPER ? (a[pos] <= key) : (a[pos] < key));
    return lo + __popc(__ballot_sync(0xffffffffu, below));
}

// rank of every detection inside its frame by (x, index): sorted copies of the positions + permutation.
// One block = 256 consecutive detections. The x coordinates of the frames the block touches (only the first and the
// last can stick out of it) are staged in shared memory; a detection's rank is the number of earlier entries of its
// frame with x <= its own plus the number of later ones with x < its own (stable), counted from shared memory.
constexpr int FS_TPB = 256;
constexpr int FS_CAP = 2048;  // staged x values; larger frame spans are read in place

__global__ void __launch_bounds__(FS_TPB) k_frame_sort(int64_t n, int dim, const int32_t *__restrict__ tkey,
                                                       const double *__restrict__ pos, double *__restrict__ spos,
                                                       int32_t *__restrict__ sperm, int64_t row_begin, int64_t row_end) {
    // rows [row_begin, row_end) of arrays of n detections (whole graphs: no frame straddles the range)
    __shared__ int s_key[FS_TPB];
    __shared__ int64_t s_edge[2];  // start of the block's first frame, end of its last frame
    __shared__ double s_x[FS_CAP];
    const int64_t b0 = row_begin + (int64_t)blockIdx.x * FS_TPB;
    const int64_t b1 = (b0 + FS_TPB < row_end) ? b0 + FS_TPB : row_end;
    const int64_t i = b0 + threadIdx.x;
    const bool active = i < row_end;
    const int k = active ? tkey[i] : 0;
    s_key[threadIdx.x] = k;
    if (threadIdx.x < 32) {
        const int64_t lo = warp_bound_i32<false>(tkey, row_begin, b0 + 1, tkey[b0]);
        if (threadIdx.x == 0) s_edge[0] = lo;
    } else if (threadIdx.x < 64) {
        const int64_t hi = warp_bound_i32<true>(tkey, b1 - 1, row_end, tkey[b1 - 1]);
        if (threadIdx.x == 32) s_edge[1] = hi;
    }
    __syncthreads();
    const int64_t e0 = s_edge[0], e1 = s_edge[1];
    const bool staged = e1 - e0 <= FS_CAP;
    if (staged)
        for (int c = threadIdx.x; c < (int)(e1 - e0); c += FS_TPB) s_x[c] = pos[e0 + c];
    __syncthreads();
    if (!active) return;
    const int nb = (int)(b1 - b0);
    // the detection's frame inside the block, [l, h): the keys are sorted, two binary searches over the block's keys
    int l = 0, h = nb;
    {
        int a = 0, z = threadIdx.x;  // first index with key == k lies in [0, tid]
        while (a < z) {
            const int mid = (a + z) >> 1;
            if (s_key[mid] < k)
                a = mid + 1;
            else
                z = mid;
        }
        l = a;
        a = threadIdx.x + 1;
        z = nb;  // first index with key > k lies in [tid + 1, nb]
        while (a < z) {
            const int mid = (a + z) >> 1;
            if (s_key[mid] <= k)
                a = mid + 1;
            else
                z = mid;
        }
        h = a;
    }
    const int64_t lo = (l == 0 && s_key[0] == k) ? e0 : b0 + l;
    const int64_t hi = (h == nb && s_key[nb - 1] == k) ? e1 : b0 + h;
    int rank = 0;
    if (staged) {
        // one loop over the whole frame for all of its detections (same trip count in every lane of a frame):
        // strictly smaller entries, and how many are equal; ties (rare) are ranked by index in a second pass
        const int a = (int)(lo - e0), me = (int)(i - e0), z = (int)(hi - e0);
        const double x = s_x[me];
        int eq = 0;
#pragma unroll 4
        for (int j = a; j < z; j++) {
            const double xj = s_x[j];
            rank += (xj < x) ? 1 : 0;
            eq += (xj == x) ? 1 : 0;
        }
        if (eq > 1)
            for (int j = a; j < me; j++) rank += (s_x[j] == x) ? 1 : 0;
    } else {
        const double x = pos[i];
        for (int64_t j = lo; j < hi; j++) {
            const double xj = pos[j];
            rank += (xj < x || (xj == x && j < i)) ? 1 : 0;
        }
    }
    const int64_t o = lo + rank;
    sperm[o] = (int32_t)i;
    for (int d = 0; d < dim; d++) spos[(size_t)d * n + o] = pos[(size_t)d * n + i];
}

// the arcs of one row: x-window search in each of the <= max_jump frames it may link to. cx/cy/cz/cperm/ckey = the
// tile's candidate range (shared memory copies when STAGED, the global arrays otherwise), m entries.
template <int DIR, bool FILL, bool STAGED>
__device__ __forceinline__ int row_arcs(const PrunedParams &p, const int64_t r, const int64_t kmin, const bool tabled,
                                        const int *s_seg_lo, const int *s_seg_hi, const double *cx, const double *cy,
                                        const double *cz, const int32_t *cperm, const int32_t *ckey, const int m) {
    const int64_t n = p.n;
    const int J = p.max_jump;
    const bool has_z = p.dim == 3;
    const int key_r = p.tkey[r];
    const double xr = p.pos[r], yr = p.pos[n + r];
    const double zr = has_z ? p.pos[2 * n + r] : 0.0;
    const double xlo = xr - p.x_window, xhi = xr + p.x_window;
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
            a = (int)bound_i32<false>(ckey, 0, m, kt);
            b = (int)bound_i32<true>(ckey, a, m, kt);
        }
        if (b <= a) continue;
        // first candidate with x >= xr - window
        int l = a, h = b;
        while (l < h) {
            const int mid = (l + h) >> 1;
            if (cx[mid] < xlo)
                l = mid + 1;
            else
                h = mid;
        }
        // The survivors of this frame come out in x order; the reference emits them in detection order
        // (mot3d/mot3d.py:111-127). Up to 4 are sorted in registers, a longer run falls back to an insertion into the
        // output arrays.
        int nf = 0, o0 = 0, o1 = 0, o2 = 0, o3 = 0;
        double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
        for (int c = l; c < b; c++) {
            const double xc = cx[c];
            if (xc > xhi) break;
            const double dx = xr - xc;
            const double dy = yr - cy[c];
            double s = dx * dx;  // sum([dx^2, dy^2, dz^2]) left to right, as the reference's euclidean()
            s = s + dy * dy;
            if (has_z) {
                const double dz = zr - cz[c];
                s = s + dz * dz;
            }
            if (s <= p.sq_dist_max) {
                if (FILL) {
                    const int other = cperm[c];
                    if (nf == 0) {
                        o0 = other;
                        s0 = s;
                    } else if (nf == 1) {
                        o1 = other;
                        s1 = s;
                    } else if (nf == 2) {
                        o2 = other;
                        s2 = s;
                    } else if (nf == 3) {
                        o3 = other;
                        s3 = s;
                    } else {
                        if (nf == 4) {  // spill the four held back, in x order; the insertion below sorts them
                            const int oo[4] = {o0, o1, o2, o3};
                            const double ss[4] = {s0, s1, s2, s3};
                            for (int t = 0; t < 4; t++) {
                                int64_t e = out_base + cnt + t;
                                if (e < p.edge_capacity) {
                                    while (e > out_base + cnt && p.col[e - 1] > oo[t]) {
                                        p.col[e] = p.col[e - 1];
                                        p.cost[e] = p.cost[e - 1];
                                        e--;
                                    }
                                    p.col[e] = oo[t];
                                    p.cost[e] = __double_as_longlong(ss[t]);
                                } else if (p.overflow) {
                                    *p.overflow = 1;
                                }
                            }
                        }
                        int64_t e = out_base + cnt + nf;
                        if (e < p.edge_capacity) {
                            while (e > out_base + cnt && p.col[e - 1] > other) {
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
                }
                nf++;
            }
        }
        if (FILL && nf > 0 && nf <= 4) {
            // sorting network on (detection, squared distance) pairs; unused slots sort last
            if (nf < 2) o1 = 0x7fffffff;
            if (nf < 3) o2 = 0x7fffffff;
            if (nf < 4) o3 = 0x7fffffff;
#define M3_CSWAP(oa, sa, ob, sb)      \
    if (oa > ob) {                    \
        const int to_ = oa;           \
        oa = ob;                      \
        ob = to_;                     \
        const double ts_ = sa;        \
        sa = sb;                      \
        sb = ts_;                     \
    }
            if (nf > 1) {
                M3_CSWAP(o0, s0, o1, s1);
                if (nf > 2) {
                    M3_CSWAP(o2, s2, o3, s3);
                    M3_CSWAP(o0, s0, o2, s2);
                    M3_CSWAP(o1, s1, o3, s3);
                    M3_CSWAP(o1, s1, o2, s2);
                }
            }
#undef M3_CSWAP
            const int64_t e = out_base + cnt;
            if (e + nf <= p.edge_capacity) {
                p.col[e] = o0;
                p.cost[e] = __double_as_longlong(s0);
                if (nf > 1) {
                    p.col[e + 1] = o1;
                    p.cost[e + 1] = __double_as_longlong(s1);
                }
                if (nf > 2) {
                    p.col[e + 2] = o2;
                    p.cost[e + 2] = __double_as_longlong(s2);
                }
                if (nf > 3) {
                    p.col[e + 3] = o3;
                    p.cost[e + 3] = __double_as_longlong(s3);
                }
            } else if (p.overflow) {
                *p.overflow = 1;
            }
        }
        cnt += nf;
    }
    return cnt;
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

    // candidate index range of the tile (the sorted copies permute detections inside frames only): two warps search
    // the two ends; the keys are sorted, so the range lies before (DIR_IN) / after (DIR_OUT) the tile
    const int64_t k_first = p.tkey[r0], k_last = p.tkey[r1 - 1];
    const int64_t kmin = (DIR == MOT3D_DIR_OUT) ? k_first + 1 : k_first - J;
    const int64_t kmax = (DIR == MOT3D_DIR_OUT) ? k_last + J : k_last - 1;
    if (threadIdx.x < 32) {
        const int64_t v = warp_bound_i32<false>(p.tkey, (DIR == MOT3D_DIR_OUT) ? r0 : 0, (DIR == MOT3D_DIR_OUT) ? n : r1, kmin);
        if (threadIdx.x == 0) {
            s_range[0] = v;
            mbar_init(&s_bar, 1);
            fence_mbar_init();
        }
    } else if (threadIdx.x < 64) {
        const int64_t v = warp_bound_i32<true>(p.tkey, (DIR == MOT3D_DIR_OUT) ? r0 : 0, (DIR == MOT3D_DIR_OUT) ? n : r1, kmax);
        if (threadIdx.x == 32) s_range[1] = v;
    }
    for (int t = threadIdx.x; t < EP_TABLE; t += EP_TPB) {
        s_seg_lo[t] = 0;
        s_seg_hi[t] = 0;
    }
    __syncthreads();
    const int64_t lo = s_range[0], hi = s_range[1] > s_range[0] ? s_range[1] : s_range[0];
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
    int cnt;
    if (staged)
        cnt = row_arcs<DIR, FILL, true>(p, r, kmin, tabled, s_seg_lo, s_seg_hi, s_x, s_y, s_z, s_perm, s_key, m);
    else
        cnt = row_arcs<DIR, FILL, false>(p, r, kmin, tabled, s_seg_lo, s_seg_hi, p.spos + lo, p.spos + n + lo,
                                         p.spos + 2 * n + lo, p.sperm + lo, p.tkey + lo, m);
    if (!FILL) p.row_count[r] = cnt;
}

}  // namespace m3

using namespace m3;

// internal entry points used by mot3d_edge_count / mot3d_edge_fill (edge_build.cu) when sorted copies are supplied
namespace m3 {

int32_t launch_frame_sort_range(const mot3d_dets *d, int64_t row_begin, int64_t row_end, cudaStream_t st) {
    if (row_end <= row_begin) return MOT3D_OK;
    k_frame_sort<<<(unsigned)((row_end - row_begin + FS_TPB - 1) / FS_TPB), FS_TPB, 0, st>>>(
        d->n, d->dim, d->tkey, d->pos, d->sorted_pos, d->sorted_perm, row_begin, row_end);
    M3_LAUNCH_CHECK("k_frame_sort");
    return MOT3D_OK;
}

int32_t launch_frame_sort(const mot3d_dets *d, cudaStream_t st) { return launch_frame_sort_range(d, 0, d->n, st); }

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
