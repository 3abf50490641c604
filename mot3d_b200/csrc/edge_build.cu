// Graph build kernels: time keys, per-detection arc costs, the windowed all-pairs transition kernel
// (count pass / fill pass), the appearance (colour histogram) pass and the row scan.
//
// Reference semantics (cvlab-epfl/mot3d):
//   build_graph hot loop            mot3d/mot3d.py:106-141   all ordered pairs 0 < frame[j]-frame[i] <= max_jump
//   weight_distance_detections_2d   mot3d/weight_functions.py:13-35 (3D: :43-75)
//   weight_confidence_detections    mot3d/weight_functions.py:37-38
//   bbox_size_similarity            mot3d/distance_functions.py:36-45
//   color_histogram_similarity      mot3d/distance_functions.py:21-34
//   quantisation                    mot3d/mot3d.py:224 ("{:0.7f}")
// All float math is FP64 in the reference's operation order (this file is compiled with -fmad=false; the
// only fused operations are the explicit fma() calls).
#include "common.cuh"

namespace m3 {

constexpr int EB_TPB = 128;    // one thread per CSR row, 4 warps
constexpr int EB_CHUNK = 1024; // candidates staged in shared memory per step

struct EdgeParams {
    int64_t n;
    int32_t dim;
    int32_t max_jump;
    int32_t use_bbox;
    int32_t use_hist;
    const int32_t *tkey;
    const double *pos;
    const double *bbox;
    double sq_dist_max;
    double sigma_distance_sq;
    double sigma_bbox_sq;
    JumpTable jt;  // travels in the kernel parameter block (constant bank), indexed by jump-1
    // outputs
    int32_t *row_count;
    const int32_t *row_ptr;
    int64_t edge_capacity;
    int32_t *col;
    int64_t *cost;
    double *weight;
    int32_t *overflow;
};

// first index in [0,n) whose key is > key (UPPER) or >= key (!UPPER); whole warp cooperates (32-ary search).
template <bool UPPER>
__device__ __forceinline__ int64_t warp_bound(const int32_t *__restrict__ a, int64_t n, int64_t key) {
    int64_t lo = 0, hi = n;
    const int l = lane_id();
    while (lo < hi) {
        int64_t span = hi - lo;
        int64_t step = span / 32;
        if (step < 1) step = 1;
        int64_t idx = lo + (int64_t)l * step;
        bool before = false;
        if (idx < hi) {
            int64_t v = a[idx];
            before = UPPER ? (v <= key) : (v < key);
        }
        unsigned m = __ballot_sync(0xffffffffu, before);
        int cnt = __popc(m);
        if (cnt == 0) {
            hi = lo;
        } else {
            int64_t new_lo = lo + (int64_t)(cnt - 1) * step + 1;
            int64_t cand_hi = lo + (int64_t)cnt * step;
            if (cnt < 32 && cand_hi < hi) hi = cand_hi;
            lo = new_lo;
        }
    }
    return lo;
}

template <int DIR, bool FILL>
__global__ void __launch_bounds__(EB_TPB) k_edge_build(const __grid_constant__ EdgeParams p) {
    __shared__ __align__(16) double s_x[EB_CHUNK + 2];
    __shared__ __align__(16) double s_y[EB_CHUNK + 2];
    __shared__ __align__(16) double s_z[EB_CHUNK + 2];
    __shared__ __align__(16) int32_t s_k[EB_CHUNK + 4];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ int64_t s_range[2];

    const int64_t n = p.n;
    const int64_t r0 = (int64_t)blockIdx.x * EB_TPB;
    const int64_t r1 = (r0 + EB_TPB < n) ? r0 + EB_TPB : n;
    const int64_t r = r0 + threadIdx.x;
    const bool active = r < n;
    const int J = p.max_jump;
    const bool has_z = p.dim == 3;

    if (threadIdx.x == 0) {
        mbar_init(&s_bar, 1);
        fence_mbar_init();
    }
    // candidate index range of the whole tile (keys are non-decreasing over the batch)
    if (warp_id() == 0) {
        int64_t k_first = p.tkey[r0], k_last = p.tkey[r1 - 1];
        int64_t lo, hi;
        if (DIR == MOT3D_DIR_OUT) {
            lo = warp_bound<true>(p.tkey, n, k_first);       // first key > k_first
            hi = warp_bound<true>(p.tkey, n, k_last + J);    // first key > k_last + J
        } else {
            lo = warp_bound<false>(p.tkey, n, k_first - J);  // first key >= k_first - J
            hi = warp_bound<false>(p.tkey, n, k_last);       // first key >= k_last
        }
        if (lane_id() == 0) {
            s_range[0] = lo;
            s_range[1] = hi;
        }
    }
    int32_t key_r = 0;
    double xr = 0, yr = 0, zr = 0;
    if (active) {
        key_r = p.tkey[r];
        xr = p.pos[r];
        yr = p.pos[n + r];
        if (has_z) zr = p.pos[2 * n + r];
    }
    __syncthreads();
    const int64_t lo = s_range[0], hi = s_range[1];
    int cnt = 0;
    int64_t out_base = 0;
    if (FILL && active) out_base = p.row_ptr[r];
    uint32_t phase = 0;

    for (int64_t c0 = lo; c0 < hi; c0 += EB_CHUNK) {
        const int m = (int)((hi - c0 < EB_CHUNK) ? (hi - c0) : EB_CHUNK);
        uint32_t bulk = 0;
        const int shx = stage_slice(p.pos, s_x, c0, m, &s_bar, &bulk);
        const int shy = stage_slice(p.pos + n, s_y, c0, m, &s_bar, &bulk);
        int shz = 0;
        if (has_z) shz = stage_slice(p.pos + 2 * n, s_z, c0, m, &s_bar, &bulk);
        const int shk = stage_slice(p.tkey, s_k, c0, m, &s_bar, &bulk);
        if (threadIdx.x == 0) mbar_expect_tx(&s_bar, bulk);
        __syncthreads();  // ragged head/tail stores visible
        mbar_wait(&s_bar, phase);
        phase ^= 1;

        if (active) {
#pragma unroll 4
            for (int c = 0; c < m; c++) {
                const int dk = (DIR == MOT3D_DIR_OUT) ? (s_k[shk + c] - key_r) : (key_r - s_k[shk + c]);
                if ((unsigned)(dk - 1) < (unsigned)J) {
                    const double dx = xr - s_x[shx + c];
                    const double dy = yr - s_y[shy + c];
                    double s = dx * dx;  // sum([dx^2, dy^2, dz^2]) left to right
                    s = s + dy * dy;
                    if (has_z) {
                        const double dz = zr - s_z[shz + c];
                        s = s + dz * dz;
                    }
                    if (s <= p.sq_dist_max) {  // <=> not (sqrt(s) > max_distance)
                        if (FILL) {
                            // only the column and the squared distance are stored here (two plain stores); the
                            // weight needs sqrt/exp/div and is computed by k_edge_cost for the ~1 % of pairs that
                            // survive, instead of dragging the whole warp through it at every hit
                            const int64_t e = out_base + cnt;
                            if (e < p.edge_capacity) {
                                p.col[e] = (int32_t)(c0 + c);
                                p.cost[e] = __double_as_longlong(s);
                            } else if (p.overflow) {
                                *p.overflow = 1;
                            }
                        }
                        cnt++;
                    }
                }
            }
        }
        __syncthreads();  // everyone done with this chunk before it is overwritten
    }
    if (!FILL && active) p.row_count[r] = cnt;
}

// Second half of the fill: one thread per row turns (column, squared distance) into the arc weight
//   w = (-exp(-(jump-1)*sigma_jump)) * ((exp(-dist^2/sigma_d^2) [* w_hist]) * w_bbox)
// (reference mot3d/weight_functions.py:13-35) and its exact quantisation (mot3d/mot3d.py:224).
// With use_hist the distance factor alone is stored in weight[]: k_edge_appearance finishes (the reference
// multiplies the histogram factor in before the bbox factor).
template <int DIR>
__global__ void __launch_bounds__(128) k_edge_cost(const __grid_constant__ EdgeParams p) {
    // one warp per 32 consecutive rows: their arcs are one contiguous range, walked with coalesced accesses; the
    // row of an arc is found by a binary search over the 32 row starts held one per lane
    const int64_t n = p.n;
    const int64_t r0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + warp_id()) * 32;
    if (r0 >= n) return;
    const int lane = lane_id();
    const int64_t r_lane = (r0 + lane < n) ? r0 + lane : n - 1;
    const int rp = p.row_ptr[(r0 + lane <= n) ? r0 + lane : n];  // lanes past the end repeat the final offset
    const int key_lane = p.tkey[r_lane];
    int64_t e_begin = __shfl_sync(0xffffffffu, rp, 0);
    int64_t e_end = p.row_ptr[(r0 + 32 < n) ? r0 + 32 : n];
    if (e_end > p.edge_capacity) e_end = p.edge_capacity;
    for (int64_t e0 = e_begin; e0 < e_end; e0 += 32) {
        const int64_t e = e0 + lane;
        const bool act = e < e_end;
        // largest lane l with rp[l] <= e
        int lo = 0;
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
            const int cand = lo + step;
            const int v = __shfl_sync(0xffffffffu, rp, cand & 31);
            if (cand < 32 && v <= (int)e) lo = cand;
        }
        const int key_r = __shfl_sync(0xffffffffu, key_lane, lo);
        if (!act) continue;
        const int64_t r = r0 + lo;
        const int64_t other = p.col[e];
        const double s = __longlong_as_double(p.cost[e]);
        const int dk = (DIR == MOT3D_DIR_OUT) ? (p.tkey[other] - key_r) : (key_r - p.tkey[other]);
        double prod = distance_factor(s, p.sigma_distance_sq);
        double w;
        if (p.use_hist) {
            w = prod;
        } else {
            if (p.use_bbox) {
                const int64_t a = (DIR == MOT3D_DIR_OUT) ? r : other;
                const int64_t b = (DIR == MOT3D_DIR_OUT) ? other : r;
                const double bss = 1.0 - bbox_similarity(p.bbox, n, a, b);
                prod = prod * exp(-(bss * bss) / p.sigma_bbox_sq);
            }
            w = p.jt.v[dk - 1] * prod;
            p.cost[e] = quantise_1e7(w);
        }
        if (p.weight) p.weight[e] = w;
    }
}

// One warp per row: multiply the colour-histogram factor in, then bbox and jump factors, then quantise.
// NCC per channel = sum(h1*h2) / sqrt(sum(h1^2) * sum(h2^2)), averaged over channels (unit weights).
template <int DIR>
__global__ void __launch_bounds__(128) k_edge_appearance(int64_t n, const int32_t *__restrict__ tkey,
                                                         const double *__restrict__ hist, int C, int B,
                                                         const double *__restrict__ bbox, int use_bbox,
                                                         double sigma_hist_sq, double sigma_bbox_sq,
                                                         const __grid_constant__ JumpTable jt,
                                                         const int32_t *__restrict__ row_ptr,
                                                         const int32_t *__restrict__ col, double *__restrict__ weight,
                                                         int64_t *__restrict__ cost, int64_t arc_cap) {
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp_id();
    if (r >= n) return;
    const int l = lane_id();
    // (rows reaching beyond the arc arrays: the build ran out of capacity and has reported it)
    const int e0 = (int)min((int64_t)row_ptr[r], arc_cap), e1 = (int)min((int64_t)row_ptr[r + 1], arc_cap);
    const double *hr = hist + r * (int64_t)C * B;
    for (int e = e0; e < e1; e++) {
        const int64_t o = col[e];
        const double *ho = hist + o * (int64_t)C * B;
        double scs = 0.0;
        for (int c = 0; c < C; c++) {
            double ab = 0, aa = 0, bb = 0;
            for (int b = l; b < B; b += 32) {
                double u = hr[c * B + b], v = ho[c * B + b];
                ab += u * v;
                aa += u * u;
                bb += v * v;
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                ab += __shfl_xor_sync(0xffffffffu, ab, d);
                aa += __shfl_xor_sync(0xffffffffu, aa, d);
                bb += __shfl_xor_sync(0xffffffffu, bb, d);
            }
            scs = scs + ab / sqrt(aa * bb);
        }
        if (l == 0) {
            const double chs = 1.0 - scs / (double)C;
            double prod = weight[e] * exp(-(chs * chs) / sigma_hist_sq);
            if (use_bbox) {
                const int64_t a = (DIR == MOT3D_DIR_OUT) ? r : o;
                const int64_t b = (DIR == MOT3D_DIR_OUT) ? o : r;
                const double bss = 1.0 - bbox_similarity(bbox, n, a, b);
                prod = prod * exp(-(bss * bss) / sigma_bbox_sq);
            }
            const int dk = (DIR == MOT3D_DIR_OUT) ? (tkey[o] - tkey[r]) : (tkey[r] - tkey[o]);
            const double w = jt.v[dk - 1] * prod;
            weight[e] = w;
            cost[e] = quantise_1e7(w);
        }
    }
}

__global__ void k_node_costs(int64_t n, const double *__restrict__ conf, const uint8_t *__restrict__ flags,
                             double mul, double bias, int64_t entry_cost, int64_t exit_cost,
                             int64_t *__restrict__ en, int64_t *__restrict__ cn, int64_t *__restrict__ ex) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double w = -(conf[i] * mul + bias);  // -(confidence*mul + bias), no fma (compiled with -fmad=false)
    cn[i] = quantise_1e7(w);
    const int f = flags ? flags[i] : (MOT3D_FLAG_ENTRY | MOT3D_FLAG_EXIT);
    en[i] = (f & MOT3D_FLAG_ENTRY) ? entry_cost : INF;
    ex[i] = (f & MOT3D_FLAG_EXIT) ? exit_cost : INF;
}

// ---------------------------------------------------------------------------------------- time keys
__global__ void k_graph_base(const int32_t *__restrict__ frame, const int32_t *__restrict__ graph_ptr, int n_graphs,
                             int max_jump, int64_t *__restrict__ base) {
    // single block: base[g] = sum_{h<g} (span_h + max_jump + 1), span = last frame - first frame (0 if empty)
    __shared__ int64_t s_carry;
    __shared__ int64_t s_warp[33];
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int g0 = 0; g0 < n_graphs; g0 += blockDim.x) {
        int g = g0 + threadIdx.x;
        int64_t v = 0;
        if (g < n_graphs) {
            int a = graph_ptr[g], b = graph_ptr[g + 1];
            int64_t span = (b > a) ? ((int64_t)frame[b - 1] - (int64_t)frame[a]) : 0;
            v = (span > 0 ? span : 0) + max_jump + 1;  // (negative: unsorted frames, reported by k_make_keys)
        }
        int64_t incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int64_t t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane_id() >= d) incl += t;
        }
        if (lane_id() == 31) s_warp[warp_id()] = incl;
        __syncthreads();
        if (warp_id() == 0) {
            int nw = (blockDim.x + 31) >> 5;
            int64_t s = (lane_id() < nw) ? s_warp[lane_id()] : 0;
            int64_t si = s;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                int64_t t = __shfl_up_sync(0xffffffffu, si, d);
                if (lane_id() >= d) si += t;
            }
            s_warp[lane_id()] = si - s;
            if (lane_id() == 31) s_warp[32] = si;
        }
        __syncthreads();
        if (g < n_graphs) base[g] = s_carry + s_warp[warp_id()] + incl - v;
        __syncthreads();
        if (threadIdx.x == 0) s_carry += s_warp[32];
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) k_make_keys(const int32_t *__restrict__ frame,
                                                   const int32_t *__restrict__ graph_ptr, int n_graphs,
                                                   const int64_t *__restrict__ base, int32_t *__restrict__ tkey,
                                                   int32_t *__restrict__ err) {
    // one block per graph (grid-stride over graphs): a 5 000-detection window is 20 coalesced steps. Every build path
    // and the solver's layering rely on keys that are non-decreasing and fit 32 bits with room for a jump: both are
    // checked on the way (err: MOT3D_ERR_UNSORTED / MOT3D_ERR_KEY_RANGE, reported by the caller of the build).
    // A graph that fails gets harmless keys (all equal / clamped: monotone, no arcs worth anything), so that the kernels
    // behind this one stay inside their arrays whatever the caller does with the flag.
    int bad = 0;
    const int64_t key_max = (int64_t)0x7fffffff - MOT3D_MAX_JUMP - 1;
    for (int g = blockIdx.x; g < n_graphs; g += gridDim.x) {
        const int a = graph_ptr[g], b = graph_ptr[g + 1];
        if (b <= a) continue;
        int unsorted = 0;
        for (int i = a + 1 + threadIdx.x; i < b; i += blockDim.x) unsorted |= frame[i - 1] > frame[i] ? 1 : 0;
        unsorted = __syncthreads_or(unsorted);
        if (unsorted) bad |= MOT3D_ERR_UNSORTED;
        const int64_t shift = base[g] - (int64_t)frame[a];
        for (int i = a + threadIdx.x; i < b; i += blockDim.x) {
            int64_t key = unsorted ? base[g] : (int64_t)frame[i] + shift;
            if (key < 0 || key > key_max) {
                bad |= MOT3D_ERR_KEY_RANGE;
                key = key < 0 ? 0 : key_max;
            }
            tkey[i] = (int32_t)key;
        }
    }
    if (bad && err) atomicOr(err, bad);
}

// ---------------------------------------------------------------------------------------- exclusive scan (3 phases)
constexpr int SCAN_TPB = 256;
constexpr int SCAN_ITEMS = 16;  // per thread
constexpr int SCAN_TILE = SCAN_TPB * SCAN_ITEMS;

__global__ void __launch_bounds__(SCAN_TPB) k_scan_partials(const int32_t *__restrict__ in, int64_t n,
                                                            int64_t *__restrict__ partial) {
    __shared__ int64_t s[SCAN_TPB / 32];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    int64_t acc = 0;
    for (int k = 0; k < SCAN_ITEMS; k++) {
        int64_t i = base + (int64_t)k * SCAN_TPB + threadIdx.x;
        if (i < n) acc += in[i];
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if (lane_id() == 0) s[warp_id()] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        int64_t t = 0;
        for (int w = 0; w < SCAN_TPB / 32; w++) t += s[w];
        partial[blockIdx.x] = t;
    }
}

__global__ void k_scan_offsets(int64_t *__restrict__ partial, int64_t n_tiles) {
    // single block, in-place exclusive scan of the tile sums
    __shared__ int64_t s_warp[33];
    __shared__ int64_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int64_t t0 = 0; t0 < n_tiles; t0 += blockDim.x) {
        int64_t t = t0 + threadIdx.x;
        int64_t v = (t < n_tiles) ? partial[t] : 0;
        int64_t incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int64_t u = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane_id() >= d) incl += u;
        }
        if (lane_id() == 31) s_warp[warp_id()] = incl;
        __syncthreads();
        if (warp_id() == 0) {
            int nw = (blockDim.x + 31) >> 5;
            int64_t s = (lane_id() < nw) ? s_warp[lane_id()] : 0;
            int64_t si = s;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                int64_t u = __shfl_up_sync(0xffffffffu, si, d);
                if (lane_id() >= d) si += u;
            }
            s_warp[lane_id()] = si - s;
            if (lane_id() == 31) s_warp[32] = si;
        }
        __syncthreads();
        if (t < n_tiles) partial[t] = s_carry + s_warp[warp_id()] + incl - v;
        __syncthreads();
        if (threadIdx.x == 0) s_carry += s_warp[32];
        __syncthreads();
    }
}

__global__ void __launch_bounds__(SCAN_TPB) k_scan_apply(const int32_t *__restrict__ in, int64_t n,
                                                         const int64_t *__restrict__ partial,
                                                         int32_t *__restrict__ out) {
    // each thread owns SCAN_ITEMS consecutive items of the tile
    __shared__ int s_scratch[33];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        int64_t i = base + k;
        v[k] = (i < n) ? in[i] : 0;
        sum += v[k];
    }
    int total;
    int excl = block_excl_scan(sum, s_scratch, &total);
    int64_t run = partial[blockIdx.x] + excl;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        int64_t i = base + k;
        if (i < n) out[i] = (int32_t)run;
        run += v[k];
        if (i == n - 1) out[n] = (int32_t)run;
    }
    if (n == 0 && blockIdx.x == 0 && threadIdx.x == 0) out[0] = 0;
}

// Multi-view appearance of 3-D detections (weight_distance_detections_3d, mot3d/weight_functions.py:55-71):
//   chs_v = 1 - color_histogram_similarity(d1.detections_2d[v].color_histogram, d2.detections_2d[v].color_histogram)
//   bss_v = 1 - bbox_size_similarity(d1.detections_2d[v].bbox, d2.detections_2d[v].bbox)
// for every view v of d1 (dict order) that d2 also has; the factors are exp(-mean(chs)^2/sigma_c^2) and
// exp(-mean(bss)^2/sigma_b^2), multiplied into the distance factor in that order (np.prod), then the jump factor.
// One warp per row; d1 = arc tail, d2 = arc head. np.mean of < 8 values adds them left to right.
template <int DIR>
__global__ void __launch_bounds__(128) k_edge_appearance_views(
    int64_t n, const int32_t *__restrict__ tkey, const int32_t *__restrict__ view_ptr,
    const int32_t *__restrict__ view_id, const double *__restrict__ vbbox, int64_t n_entries,
    const double *__restrict__ vhist, int C, int B, int use_hist, int use_bbox, double sigma_hist_sq,
    double sigma_bbox_sq, const __grid_constant__ JumpTable jt, const int32_t *__restrict__ row_ptr,
    const int32_t *__restrict__ col, double *__restrict__ weight, int64_t *__restrict__ cost,
    int32_t *__restrict__ error_flag, int64_t arc_cap) {
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp_id();
    if (r >= n) return;
    const int l = lane_id();
    const int e0 = (int)min((int64_t)row_ptr[r], arc_cap), e1 = (int)min((int64_t)row_ptr[r + 1], arc_cap);
    for (int e = e0; e < e1; e++) {
        const int64_t o = col[e];
        const int64_t a = (DIR == MOT3D_DIR_OUT) ? r : o;  // d1: the earlier detection
        const int64_t b = (DIR == MOT3D_DIR_OUT) ? o : r;  // d2
        const int a0 = view_ptr[a], a1 = view_ptr[a + 1], b0 = view_ptr[b], b1 = view_ptr[b + 1];
        double sum_c = 0.0, sum_b = 0.0;
        int shared = 0;
        for (int u = a0; u < a1; u++) {
            const int id = view_id[u];
            int v = -1;
            for (int t = b0; t < b1; t++)
                if (view_id[t] == id) {
                    v = t;
                    break;
                }
            if (v < 0) continue;
            shared++;
            if (use_hist) {
                const double *h1 = vhist + (int64_t)u * C * B;
                const double *h2 = vhist + (int64_t)v * C * B;
                double scs = 0.0;
                for (int c = 0; c < C; c++) {
                    double ab = 0, aa = 0, bb = 0;
                    for (int k = l; k < B; k += 32) {
                        const double x = h1[c * B + k], y = h2[c * B + k];
                        ab += x * y;
                        aa += x * x;
                        bb += y * y;
                    }
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) {
                        ab += __shfl_xor_sync(0xffffffffu, ab, d);
                        aa += __shfl_xor_sync(0xffffffffu, aa, d);
                        bb += __shfl_xor_sync(0xffffffffu, bb, d);
                    }
                    scs = scs + ab / sqrt(aa * bb);
                }
                sum_c = sum_c + (1.0 - scs / (double)C);
            }
            if (use_bbox) sum_b = sum_b + (1.0 - bbox_similarity(vbbox, n_entries, u, v));
        }
        if (l == 0) {
            double w;
            int64_t q = 0;
            if (shared == 0) {
                w = __longlong_as_double(0x7ff8000000000000ll);  // nan, as np.mean([]) gives the reference
                *error_flag = 1;
            } else {
                double prod = weight[e];
                if (use_hist) {
                    const double mc = sum_c / (double)shared;
                    prod = prod * exp(-(mc * mc) / sigma_hist_sq);
                }
                if (use_bbox) {
                    const double mb = sum_b / (double)shared;
                    prod = prod * exp(-(mb * mb) / sigma_bbox_sq);
                }
                const int dk = (DIR == MOT3D_DIR_OUT) ? (tkey[o] - tkey[r]) : (tkey[r] - tkey[o]);
                w = jt.v[dk - 1] * prod;
                q = quantise_1e7(w);
            }
            weight[e] = w;
            cost[e] = q;
        }
    }
}

static int32_t check_dets(const mot3d_dets *d, const mot3d_cost_spec *s) {
    M3_CHECK_ARG(d && s, "dets/spec is NULL");
    M3_CHECK_ARG(d->n >= 0 && d->n < ((int64_t)1 << 31), "detection count out of range");
    M3_CHECK_ARG(d->dim == 2 || d->dim == 3, "dim must be 2 or 3");
    M3_CHECK_ARG(s->max_jump >= 1 && s->max_jump <= MOT3D_MAX_JUMP, "max_jump must be in [1, MOT3D_MAX_JUMP]");
    M3_CHECK_ARG(d->n == 0 || (d->tkey && d->pos), "tkey/pos is NULL");
    M3_CHECK_ARG(!s->use_bbox || d->bbox, "use_bbox set but bbox is NULL");
    M3_CHECK_ARG(!s->use_hist || (d->hist && d->hist_channels > 0 && d->hist_bins > 0),
                 "use_hist set but hist is NULL/empty");
    return MOT3D_OK;
}

}  // namespace m3

using namespace m3;

extern "C" {

int32_t mot3d_make_keys(const int32_t *frame, const int32_t *graph_ptr, int32_t n_graphs, int32_t max_jump,
                        int32_t *tkey_out, int64_t *graph_base_ws, int32_t *error_flags, void *stream) {
    M3_CHECK_ARG(n_graphs >= 0, "n_graphs < 0");
    if (n_graphs == 0) return MOT3D_OK;
    M3_CHECK_ARG(frame && graph_ptr && tkey_out && graph_base_ws, "NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    k_graph_base<<<1, 1024, 0, st>>>(frame, graph_ptr, n_graphs, max_jump, graph_base_ws);
    M3_LAUNCH_CHECK("k_graph_base");
    int blocks = n_graphs < 148 * 64 ? n_graphs : 148 * 64;
    k_make_keys<<<blocks, 256, 0, st>>>(frame, graph_ptr, n_graphs, graph_base_ws, tkey_out, error_flags);
    M3_LAUNCH_CHECK("k_make_keys");
    return MOT3D_OK;
}

int32_t mot3d_node_costs(const mot3d_dets *dets, const mot3d_cost_spec *spec, int64_t *entry_cost_out,
                         int64_t *node_cost_out, int64_t *exit_cost_out, void *stream) {
    M3_CHECK_ARG(dets && spec, "dets/spec is NULL");
    if (dets->n == 0) return MOT3D_OK;
    M3_CHECK_ARG(dets->conf && entry_cost_out && node_cost_out && exit_cost_out, "NULL pointer");
    int64_t n = dets->n;
    k_node_costs<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        n, dets->conf, dets->flags, spec->conf_mul, spec->conf_bias, spec->entry_cost, spec->exit_cost,
        entry_cost_out, node_cost_out, exit_cost_out);
    M3_LAUNCH_CHECK("k_node_costs");
    return MOT3D_OK;
}


}  // extern "C"

namespace m3 {
// keys of graphs [0, n_graphs) of graph_ptr with their bases given (mot3d_track_batch_host computes them on the host and
// sends the batch through in pieces)
int32_t launch_make_keys_based(const int32_t *frame, const int32_t *graph_ptr, int32_t n_graphs, const int64_t *base,
                               int32_t *tkey, int32_t *err, cudaStream_t st) {
    if (n_graphs <= 0) return MOT3D_OK;
    int blocks = n_graphs < 148 * 64 ? n_graphs : 148 * 64;
    k_make_keys<<<blocks, 256, 0, st>>>(frame, graph_ptr, n_graphs, base, tkey, err);
    M3_LAUNCH_CHECK("k_make_keys");
    return MOT3D_OK;
}
int32_t launch_frame_sort(const mot3d_dets *d, cudaStream_t st);
int32_t launch_pruned(const mot3d_dets *d, const mot3d_cost_spec *s, int direction, bool fill, int32_t *row_count,
                      const int32_t *row_ptr, int64_t edge_capacity, int32_t *col, int64_t *cost, int32_t *overflow,
                      cudaStream_t st);
}  // namespace m3

extern "C" {

int32_t mot3d_frame_sort(const mot3d_dets *dets, void *stream) {
    M3_CHECK_ARG(dets, "dets is NULL");
    if (dets->n == 0) return MOT3D_OK;
    M3_CHECK_ARG(dets->tkey && dets->pos && dets->sorted_pos && dets->sorted_perm, "NULL pointer");
    M3_CHECK_ARG(dets->dim == 2 || dets->dim == 3, "dim must be 2 or 3");
    return launch_frame_sort(dets, (cudaStream_t)stream);
}

static void fill_params(EdgeParams &p, const mot3d_dets *d, const mot3d_cost_spec *s) {
    p.n = d->n;
    p.dim = d->dim;
    p.max_jump = s->max_jump;
    p.use_bbox = s->use_bbox;
    p.use_hist = (s->use_hist || s->defer_cost) ? 1 : 0;  // = "distance factor only, an appearance pass finishes"
    p.tkey = d->tkey;
    p.pos = d->pos;
    p.bbox = d->bbox;
    p.sq_dist_max = s->sq_dist_max;
    p.sigma_distance_sq = s->sigma_distance_sq;
    p.sigma_bbox_sq = s->sigma_bbox_sq;
    for (int i = 0; i < MOT3D_MAX_JUMP; i++) p.jt.v[i] = (i < s->max_jump) ? s->neg_exp_jump[i] : 0.0;
    p.row_count = nullptr;
    p.row_ptr = nullptr;
    p.edge_capacity = 0;
    p.col = nullptr;
    p.cost = nullptr;
    p.weight = nullptr;
    p.overflow = nullptr;
}

int32_t mot3d_edge_count(const mot3d_dets *dets, const mot3d_cost_spec *spec, int32_t direction,
                         int32_t *row_count_out, void *stream) {
    int32_t rc = check_dets(dets, spec);
    if (rc) return rc;
    M3_CHECK_ARG(direction == MOT3D_DIR_IN || direction == MOT3D_DIR_OUT, "bad direction");
    if (dets->n == 0) return MOT3D_OK;
    M3_CHECK_ARG(row_count_out, "row_count_out is NULL");
    EdgeParams p;
    fill_params(p, dets, spec);
    p.row_count = row_count_out;
    unsigned grid = (unsigned)((dets->n + EB_TPB - 1) / EB_TPB);
    cudaStream_t st = (cudaStream_t)stream;
    if (dets->sorted_pos && dets->sorted_perm)
        return launch_pruned(dets, spec, direction, false, row_count_out, nullptr, 0, nullptr, nullptr, nullptr, st);
    if (direction == MOT3D_DIR_IN)
        k_edge_build<MOT3D_DIR_IN, false><<<grid, EB_TPB, 0, st>>>(p);
    else
        k_edge_build<MOT3D_DIR_OUT, false><<<grid, EB_TPB, 0, st>>>(p);
    M3_LAUNCH_CHECK("k_edge_build<count>");
    return MOT3D_OK;
}

int32_t mot3d_edge_fill(const mot3d_dets *dets, const mot3d_cost_spec *spec, int32_t direction, const int32_t *row_ptr,
                        int64_t edge_capacity, int32_t *col_out, int64_t *cost_out, double *weight_out,
                        int32_t *overflow_flag, void *stream) {
    int32_t rc = check_dets(dets, spec);
    if (rc) return rc;
    M3_CHECK_ARG(direction == MOT3D_DIR_IN || direction == MOT3D_DIR_OUT, "bad direction");
    if (dets->n == 0) return MOT3D_OK;
    M3_CHECK_ARG(row_ptr && edge_capacity >= 0, "row_ptr is NULL / negative capacity");
    M3_CHECK_ARG(edge_capacity == 0 || (col_out && cost_out), "col_out/cost_out is NULL");
    M3_CHECK_ARG(!(spec->use_hist || spec->defer_cost) || weight_out,
                 "use_hist / defer_cost need weight_out (finished by mot3d_edge_appearance[_views])");
    EdgeParams p;
    fill_params(p, dets, spec);
    p.row_ptr = row_ptr;
    p.edge_capacity = edge_capacity;
    p.col = col_out;
    p.cost = cost_out;
    p.weight = weight_out;
    p.overflow = overflow_flag;
    unsigned grid = (unsigned)((dets->n + EB_TPB - 1) / EB_TPB);
    cudaStream_t st = (cudaStream_t)stream;
    if (dets->sorted_pos && dets->sorted_perm) {
        rc = launch_pruned(dets, spec, direction, true, nullptr, row_ptr, edge_capacity, col_out, cost_out,
                           overflow_flag, st);
        if (rc) return rc;
    } else {
        if (direction == MOT3D_DIR_IN)
            k_edge_build<MOT3D_DIR_IN, true><<<grid, EB_TPB, 0, st>>>(p);
        else
            k_edge_build<MOT3D_DIR_OUT, true><<<grid, EB_TPB, 0, st>>>(p);
        M3_LAUNCH_CHECK("k_edge_build<fill>");
    }
    if (edge_capacity > 0) {
        const unsigned cgrid = (unsigned)((dets->n + 127) / 128);  // 4 warps x 32 rows per block
        if (direction == MOT3D_DIR_IN)
            k_edge_cost<MOT3D_DIR_IN><<<cgrid, 128, 0, st>>>(p);
        else
            k_edge_cost<MOT3D_DIR_OUT><<<cgrid, 128, 0, st>>>(p);
        M3_LAUNCH_CHECK("k_edge_cost");
    }
    return MOT3D_OK;
}

int32_t mot3d_edge_appearance(const mot3d_dets *dets, const mot3d_cost_spec *spec, int32_t direction,
                              const int32_t *row_ptr, const int32_t *col, int64_t arc_capacity, double *weight_inout,
                              int64_t *cost_out, void *stream) {
    int32_t rc = check_dets(dets, spec);
    if (rc) return rc;
    M3_CHECK_ARG(spec->use_hist, "mot3d_edge_appearance called without use_hist");
    M3_CHECK_ARG(direction == MOT3D_DIR_IN || direction == MOT3D_DIR_OUT, "bad direction");
    if (dets->n == 0) return MOT3D_OK;
    M3_CHECK_ARG(row_ptr && col && weight_inout && cost_out, "NULL pointer");
    JumpTable jt;
    for (int i = 0; i < MOT3D_MAX_JUMP; i++) jt.v[i] = (i < spec->max_jump) ? spec->neg_exp_jump[i] : 0.0;
    unsigned grid = (unsigned)((dets->n + 3) / 4);
    cudaStream_t st = (cudaStream_t)stream;
    if (direction == MOT3D_DIR_IN)
        k_edge_appearance<MOT3D_DIR_IN><<<grid, 128, 0, st>>>(dets->n, dets->tkey, dets->hist, dets->hist_channels,
                                                              dets->hist_bins, dets->bbox, spec->use_bbox,
                                                              spec->sigma_hist_sq, spec->sigma_bbox_sq, jt, row_ptr,
                                                              col, weight_inout, cost_out, arc_capacity);
    else
        k_edge_appearance<MOT3D_DIR_OUT><<<grid, 128, 0, st>>>(dets->n, dets->tkey, dets->hist, dets->hist_channels,
                                                               dets->hist_bins, dets->bbox, spec->use_bbox,
                                                               spec->sigma_hist_sq, spec->sigma_bbox_sq, jt, row_ptr,
                                                               col, weight_inout, cost_out, arc_capacity);
    M3_LAUNCH_CHECK("k_edge_appearance");
    return MOT3D_OK;
}

int32_t mot3d_edge_appearance_views(const mot3d_dets *dets, const mot3d_views *views, const mot3d_cost_spec *spec,
                                    int32_t direction, const int32_t *row_ptr, const int32_t *col,
                                    int64_t arc_capacity, double *weight_inout, int64_t *cost_out,
                                    int32_t *error_flag, void *stream) {
    M3_CHECK_ARG(dets && views && spec, "dets/views/spec is NULL");
    M3_CHECK_ARG(direction == MOT3D_DIR_IN || direction == MOT3D_DIR_OUT, "bad direction");
    M3_CHECK_ARG(spec->max_jump >= 1 && spec->max_jump <= MOT3D_MAX_JUMP, "max_jump must be in [1, MOT3D_MAX_JUMP]");
    M3_CHECK_ARG(spec->use_hist || spec->use_bbox, "mot3d_edge_appearance_views called without use_hist / use_bbox");
    M3_CHECK_ARG(views->n == dets->n, "views->n != dets->n");
    if (dets->n == 0) return MOT3D_OK;
    M3_CHECK_ARG(dets->tkey && row_ptr && col && weight_inout && cost_out && error_flag, "NULL pointer");
    M3_CHECK_ARG(views->view_ptr && (views->n_entries == 0 || views->view_id), "view_ptr/view_id is NULL");
    M3_CHECK_ARG(!spec->use_bbox || views->n_entries == 0 || views->bbox, "use_bbox set but views->bbox is NULL");
    M3_CHECK_ARG(!spec->use_hist || views->n_entries == 0 ||
                     (views->hist && views->hist_channels > 0 && views->hist_bins > 0),
                 "use_hist set but views->hist is NULL/empty");
    JumpTable jt;
    for (int i = 0; i < MOT3D_MAX_JUMP; i++) jt.v[i] = (i < spec->max_jump) ? spec->neg_exp_jump[i] : 0.0;
    unsigned grid = (unsigned)((dets->n + 3) / 4);
    cudaStream_t st = (cudaStream_t)stream;
    if (direction == MOT3D_DIR_IN)
        k_edge_appearance_views<MOT3D_DIR_IN><<<grid, 128, 0, st>>>(
            dets->n, dets->tkey, views->view_ptr, views->view_id, views->bbox, views->n_entries, views->hist,
            views->hist_channels, views->hist_bins, spec->use_hist, spec->use_bbox, spec->sigma_hist_sq,
            spec->sigma_bbox_sq, jt, row_ptr, col, weight_inout, cost_out, error_flag, arc_capacity);
    else
        k_edge_appearance_views<MOT3D_DIR_OUT><<<grid, 128, 0, st>>>(
            dets->n, dets->tkey, views->view_ptr, views->view_id, views->bbox, views->n_entries, views->hist,
            views->hist_channels, views->hist_bins, spec->use_hist, spec->use_bbox, spec->sigma_hist_sq,
            spec->sigma_bbox_sq, jt, row_ptr, col, weight_inout, cost_out, error_flag, arc_capacity);
    M3_LAUNCH_CHECK("k_edge_appearance_views");
    return MOT3D_OK;
}

size_t mot3d_scan_workspace_bytes(int64_t n) {
    int64_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (tiles < 1) tiles = 1;
    return (size_t)tiles * sizeof(int64_t);
}

int32_t mot3d_exclusive_scan_i32(const int32_t *in, int64_t n, int32_t *out, void *ws, size_t ws_bytes, void *stream) {
    M3_CHECK_ARG(n >= 0 && out, "bad scan arguments");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) {
        M3_CUDA(cudaMemsetAsync(out, 0, sizeof(int32_t), st));
        return MOT3D_OK;
    }
    M3_CHECK_ARG(in && ws, "NULL pointer");
    int64_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (ws_bytes < (size_t)tiles * sizeof(int64_t)) {
        set_error("scan workspace too small: %zu < %zu", ws_bytes, (size_t)tiles * sizeof(int64_t));
        return MOT3D_E_CAPACITY;
    }
    int64_t *partial = (int64_t *)ws;
    k_scan_partials<<<(unsigned)tiles, SCAN_TPB, 0, st>>>(in, n, partial);
    M3_LAUNCH_CHECK("k_scan_partials");
    k_scan_offsets<<<1, 1024, 0, st>>>(partial, tiles);
    M3_LAUNCH_CHECK("k_scan_offsets");
    k_scan_apply<<<(unsigned)tiles, SCAN_TPB, 0, st>>>(in, n, partial, out);
    M3_LAUNCH_CHECK("k_scan_apply");
    return MOT3D_OK;
}

}  // extern "C"
