// Single-pass graph build: k_edge_build_fused<DIR> turns x-sorted frames into the finished transition CSR
// (row_ptr, col, quantised cost [, float64 weight]) in ONE launch. It replaces the count pass, the three scan
// kernels, the fill pass and k_edge_cost (csrc/edge_pruned.cu, csrc/edge_build.cu) on the production path; those stay
// as the cross-check (tests compare the two bit for bit) and as the brute-force fallback.
//
// Reference semantics: build_graph's transition loop, mot3d/mot3d.py:106-141; weight_distance_detections_2d/_3d,
// mot3d/weight_functions.py:13-35, 43-75; "{:0.7f}" quantisation, mot3d/mot3d.py:224.
//
// One CTA = one tile of 256 consecutive CSR rows, one thread per row:
//   1. the tile's candidate range (time keys within max_jump) is staged in shared memory through the TMA unit
//      (x-sorted positions + permutation) and a key-indexed table of frame segments is built;
//   2. pass 1: every row binary-searches its x-window in each frame it may link to, tests the candidates with the
//      reference's exact FP64 expression and keeps (window start, survivor bit mask) per (row, frame) in shared
//      memory; block scan of the row counts;
//   3. the tile's arc count goes into a decoupled look-back (one 64-bit status word per tile: flag | running
//      total): the aggregate is published at once, warp 0 walks back over its predecessors 32 tiles at a time
//      after its share of pass 2, which only needs tile-local offsets;
//   4. pass 2: the survivors are replayed from the masks into a tile-local arc list in shared memory (4 bytes per
//      arc: candidate slot, jump, row), each (row, frame) run put back into detection order;
//   5. cost pass: one thread per arc, no divergence: squared distance recomputed from the staged positions (same
//      expression, same bits), distance / bbox / jump factors, exact quantisation, and fully coalesced stores of
//      col[], cost[] (and weight[]) at the tile's global offset.
// Tiles that do not fit the staging areas (very dense frames) run the same code on the global arrays.
// The arc offset of the first tile can continue a device-side counter, so consecutive launches over row ranges
// append to one CSR (software pipeline of mot3d_track_batch_host).
#include <atomic>

#include "common.cuh"

namespace m3 {

constexpr int EF_TPB = 256;
constexpr int EF_TABLE = 448;      // distinct time keys a tile's candidate range may span and still use the table
constexpr int EF_CACHE_J = 16;     // pass-1 results are kept for max_jump <= this (4 bytes per row and frame)
constexpr uint32_t EF_REDO = 0xffffffffu;
constexpr unsigned long long EF_FLAG_AGG = 1ull << 62, EF_FLAG_PREFIX = 2ull << 62, EF_VALUE = (1ull << 62) - 1;

struct FusedParams {
    int64_t n;
    int64_t row_begin, row_end;
    int32_t dim, max_jump, use_bbox, use_hist;
    int32_t cap, acap, cache_j, reserved;
    const int32_t *tkey;
    const double *pos;
    const double *spos;
    const int32_t *sperm;
    const double *bbox;
    double sq_dist_max, x_window, sigma_distance_sq, sigma_bbox_sq;
    JumpTable jt;
    int32_t *row_ptr;
    int64_t edge_capacity;
    int32_t *col;
    int64_t *cost;
    double *weight;
    int32_t *overflow;
    unsigned long long *status;
    int32_t *ticket;
    const int64_t *arc_base_in;
    int64_t *arc_total_out;
};

template <bool UPPER>
__device__ __forceinline__ int64_t ef_warp_bound(const int32_t *a, int64_t lo, int64_t hi, int64_t key) {
    const int lane = threadIdx.x & 31;
    while (hi - lo > 32) {
        const int64_t span = hi - lo;
        const int64_t pos = lo + (span * (lane + 1)) / 33;
        const int64_t v = a[pos];
        const bool below = UPPER ? (v <= key) : (v < key);
        const unsigned m = __ballot_sync(0xffffffffu, below);
        const int c = __popc(m);
        const int64_t new_lo = c > 0 ? __shfl_sync(0xffffffffu, pos, c - 1) + 1 : lo;
        const int64_t new_hi = c < 32 ? __shfl_sync(0xffffffffu, pos, c < 32 ? c : 31) : hi;
        lo = new_lo;
        hi = new_hi;
    }
    const int64_t pos = lo + lane;
    const bool below = pos < hi && (UPPER ? (a[pos] <= key) : (a[pos] < key));
    return lo + __popc(__ballot_sync(0xffffffffu, below));
}

// The same bound when it is known to lie close to `start`: one round of 32 probes 64 entries apart, one round over the
// 64 entries between two probes - two dependent loads instead of five for a batch of millions of keys; ranges the
// probes do not cover fall back to ef_warp_bound. FORWARD: bound in [start, hi0]; else: bound in [lo0, start].
template <bool UPPER, bool FORWARD>
__device__ __forceinline__ int64_t ef_gallop(const int32_t *a, int64_t lo0, int64_t hi0, int64_t start, int64_t key) {
    const int lane = threadIdx.x & 31;
    const int64_t q = FORWARD ? start + (int64_t)(lane + 1) * 64 : start - (int64_t)(lane + 1) * 64;
    bool below;
    if (FORWARD)
        below = q < hi0 && (UPPER ? (a[q] <= key) : (a[q] < key));
    else
        below = q < lo0 || (UPPER ? (a[q] <= key) : (a[q] < key));
    const unsigned m = __ballot_sync(0xffffffffu, below);
    int64_t rlo, rhi;  // the bound lies in [rlo, rhi]
    if (FORWARD) {
        const int c = __popc(m);  // probes 0 .. c-1 are below
        if (c == 32) return ef_warp_bound<UPPER>(a, start + 32 * 64 + 1, hi0, key);
        rlo = start + (int64_t)c * 64 + (c > 0 ? 1 : 0);
        rhi = start + (int64_t)(c + 1) * 64;
        if (rhi > hi0) rhi = hi0;
    } else {
        const int f = m ? __ffs(m) - 1 : 32;  // first probe that is below
        if (f == 32) return ef_warp_bound<UPPER>(a, lo0, start - 32 * 64, key);
        rlo = start - (int64_t)(f + 1) * 64 + 1;
        if (rlo < lo0) rlo = lo0;
        rhi = start - (int64_t)f * 64;
    }
    const int64_t p0 = rlo + lane, p1 = rlo + 32 + lane;
    const bool b0 = p0 < rhi && (UPPER ? (a[p0] <= key) : (a[p0] < key));
    const bool b1 = p1 < rhi && (UPPER ? (a[p1] <= key) : (a[p1] < key));
    return rlo + __popc(__ballot_sync(0xffffffffu, b0)) + __popc(__ballot_sync(0xffffffffu, b1));
}

template <bool UPPER>
__device__ __forceinline__ int ef_bound(const int32_t *a, int lo, int hi, int key) {
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const int v = a[mid];
        if (UPPER ? (v <= key) : (v < key))
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}

// One row against the frames it may link to. FILL = false: count, and remember (window start, survivor mask) per
// frame when `cache` is set. FILL = true: write the row's arcs to pk[]/sv[] starting at `out` (replayed from the
// cache where possible), each frame's run in detection order.
template <int DIR, bool FILL, bool STAGED, bool A32>
__device__ __forceinline__ int ef_row(const FusedParams &p, const int key_r, const double xr, const double yr,
                                      const double zr, const int kmin, const bool tabled, const int *s_seg_lo,
                                      const int *s_seg_hi, const double *cx, const double *cy, const double *cz,
                                      const int32_t *cperm, const int32_t *ckey, const int m, uint32_t *cache,
                                      uint32_t *pk, double *sv, const int out) {
    const int J = p.max_jump;
    const bool has_z = p.dim == 3;
    const double xlo = xr - p.x_window, xhi = xr + p.x_window;
    const double smax = p.sq_dist_max;
    int cnt = 0;
    for (int jj = 1; jj <= J; jj++) {
        // frames in ascending time order, so that the arcs of a row ascend by detection index
        const int kt = (DIR == MOT3D_DIR_OUT) ? key_r + jj : key_r - (J + 1 - jj);
        const uint32_t jm1 = (uint32_t)((DIR == MOT3D_DIR_OUT) ? jj - 1 : J - jj);
        // arc word: candidate slot | jump-1 | (A32: row of the tile); A32 arcs carry no squared distance, the cost pass
        // recomputes it from the staged positions
        const uint32_t tag = A32 ? ((jm1 << 16) | ((uint32_t)threadIdx.x << 24)) : (jm1 << 24);
        constexpr uint32_t SLOT = A32 ? 0xffffu : 0xffffffu;
        const int start = out + cnt;
        int nf = 0, l = 0;
        uint32_t mask = 0;
        bool replay = false;
        if (FILL && cache) {
            const uint32_t cw = cache[(jj - 1) * EF_TPB + threadIdx.x];
            if (cw != EF_REDO) {
                l = (int)(cw & 0xffffu);
                mask = cw >> 16;
                replay = true;
            }
        }
        if (!replay) {
            int a, b;
            if (tabled) {
                const int t = kt - kmin;
                a = s_seg_lo[t];
                b = s_seg_hi[t];
            } else {
                a = ef_bound<false>(ckey, 0, m, kt);
                b = ef_bound<true>(ckey, a, m, kt);
            }
            if (b <= a) {
                if (!FILL && cache) cache[(jj - 1) * EF_TPB + threadIdx.x] = 0;
                continue;
            }
            // first candidate with x >= xr - window
            int len = b - a;
            l = a;
            while (len > 0) {
                const int half = len >> 1;
                const bool lt = cx[l + half] < xlo;
                l = lt ? l + half + 1 : l;
                len = lt ? len - half - 1 : half;
            }
            // The first 8 candidates of the window are tested without a branch (independent loads and FP64 chains; the
            // exact test alone decides, the x-window only bounds the search): the average window holds 6.
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int c = l + k;
                const int cc = c < b ? c : b - 1;
                const double dx = xr - cx[cc];
                const double dy = yr - cy[cc];
                double s = dx * dx;  // sum([dx^2, dy^2, dz^2]) left to right, as the reference's euclidean()
                s = s + dy * dy;
                if (has_z) {
                    const double dz = zr - cz[cc];
                    s = s + dz * dz;
                }
                // s <= smax <=> not (sqrt(s) > max_distance), see spec.sq_dist_max
                mask |= (c < b && s <= smax) ? (1u << k) : 0u;
            }
            nf = __popc(mask);
            bool big = false;
            if (l + 8 < b && cx[l + 8] <= xhi) {  // a longer window: the rest one by one
                int i = 8;
                for (int c = l + 8; c < b; c++, i++) {
                    const double xc = cx[c];
                    if (xc > xhi) break;
                    const double dx = xr - xc;
                    const double dy = yr - cy[c];
                    double s = dx * dx;
                    s = s + dy * dy;
                    if (has_z) {
                        const double dz = zr - cz[c];
                        s = s + dz * dz;
                    }
                    if (s <= smax) {
                        if (i < 16) mask |= 1u << i;
                        nf++;
                    }
                }
                big = i > 16;
            }
            if (!FILL) {
                if (cache) cache[(jj - 1) * EF_TPB + threadIdx.x] = (big || l >= 0xffff) ? EF_REDO : ((uint32_t)l | (mask << 16));
                cnt += nf;
                continue;
            }
            if (big) {  // more than 16 candidates in the window: the mask does not hold them all
                nf = 0;
                for (int c = l; c < b; c++) {
                    const double xc = cx[c];
                    if (xc > xhi) break;
                    const double dx = xr - xc;
                    const double dy = yr - cy[c];
                    double s = dx * dx;
                    s = s + dy * dy;
                    if (has_z) {
                        const double dz = zr - cz[c];
                        s = s + dz * dz;
                    }
                    if (s <= smax) {
                        pk[start + nf] = (uint32_t)c | tag;
                        if (!A32) sv[start + nf] = s;
                        nf++;
                    }
                }
                mask = 0;
            }
        }
        if (FILL) {
            if (mask) nf = 0;
            while (mask) {
                const int c = l + __ffs(mask) - 1;
                mask &= mask - 1;
                if (!A32) {
                    const double dx = xr - cx[c];
                    const double dy = yr - cy[c];
                    double s = dx * dx;
                    s = s + dy * dy;
                    if (has_z) {
                        const double dz = zr - cz[c];
                        s = s + dz * dz;
                    }
                    sv[start + nf] = s;
                }
                pk[start + nf] = (uint32_t)c | tag;
                nf++;
            }
            // the survivors of a frame come out in x order; the reference emits them in detection order
            // (mot3d/mot3d.py:111-127): a swap for the common pair, insertion sort beyond
            if (nf == 2) {
                const uint32_t w0 = pk[start], w1 = pk[start + 1];
                if (cperm[w0 & SLOT] > cperm[w1 & SLOT]) {
                    pk[start] = w1;
                    pk[start + 1] = w0;
                    if (!A32) {
                        const double s0 = sv[start], s1 = sv[start + 1];
                        sv[start] = s1;
                        sv[start + 1] = s0;
                    }
                }
            } else if (nf > 2) {
                for (int u = 1; u < nf; u++) {
                    const uint32_t w = pk[start + u];
                    const double s = A32 ? 0.0 : sv[start + u];
                    const int id = cperm[w & SLOT];
                    int v = u;
                    while (v > 0 && cperm[pk[start + v - 1] & SLOT] > id) {
                        pk[start + v] = pk[start + v - 1];
                        if (!A32) sv[start + v] = sv[start + v - 1];
                        v--;
                    }
                    pk[start + v] = w;
                    if (!A32) sv[start + v] = s;
                }
            }
        }
        cnt += nf;
    }
    return cnt;
}

__device__ __forceinline__ unsigned long long ef_ld_status(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void ef_st_status(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Exclusive prefix of tile `tile` (> 0): sum of the aggregates of its predecessors back to the first one that has a
// prefix. Whole warp 0; lane 0 publishes the tile's own prefix and leaves the exclusive one in *s_base.
__device__ __forceinline__ void ef_look_back(const FusedParams &p, const int tile, const int total, long long *s_base) {
    const int tid = threadIdx.x;
    long long excl = 0;
    int t = tile - 1;
    while (true) {
        const int idx = t - tid;
        const unsigned long long v = idx >= 0 ? ef_ld_status(p.status + idx) : EF_FLAG_PREFIX;
        const unsigned flag = (unsigned)(v >> 62);
        const unsigned bp = __ballot_sync(0xffffffffu, flag == 2);
        const unsigned bi = __ballot_sync(0xffffffffu, flag == 0);
        const int first = bp ? __ffs(bp) - 1 : 32;
        const unsigned need = first >= 31 ? 0xffffffffu : ((2u << first) - 1);
        if (bi & need) continue;  // a predecessor has not published yet
        long long c = (tid <= first) ? (long long)(v & EF_VALUE) : 0;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
        excl += c;
        if (first < 32) break;
        t -= 32;
    }
    if (tid == 0) {
        ef_st_status(p.status + tile, EF_FLAG_PREFIX | (unsigned long long)(excl + total));
        *s_base = excl;
    }
}

template <int DIR>
__global__ void __launch_bounds__(EF_TPB) k_edge_build_fused(const __grid_constant__ FusedParams p) {
    extern __shared__ __align__(16) unsigned char ef_smem[];
    __shared__ int64_t s_range[2];
    __shared__ int s_seg_lo[EF_TABLE], s_seg_hi[EF_TABLE];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ int s_scan[33];
    __shared__ int s_rowoff[EF_TPB + 1];
    __shared__ int s_tile;
    __shared__ long long s_base;

    const int tid = threadIdx.x;
    const int64_t n = p.n;
    const int J = p.max_jump;
    const bool has_z = p.dim == 3;
    // tiles are handed out in launch order, so every predecessor of a tile is resident or done: the look-back below
    // cannot wait for a block that has not started
    if (tid == 0) s_tile = atomicAdd(p.ticket, 1);
    __syncthreads();
    const int tile = s_tile;
    const int64_t r0 = p.row_begin + (int64_t)tile * EF_TPB;
    const int64_t r1 = (r0 + EF_TPB < p.row_end) ? r0 + EF_TPB : p.row_end;
    const int64_t r = r0 + tid;
    const bool active = r < r1;

    // candidate index range of the tile: the keys are sorted, so it lies before (DIR_IN) / after (DIR_OUT) the tile
    const int64_t k_first = p.tkey[r0], k_last = p.tkey[r1 - 1];
    const int64_t kmin = (DIR == MOT3D_DIR_OUT) ? k_first + 1 : k_first - J;
    const int64_t kmax = (DIR == MOT3D_DIR_OUT) ? k_last + J : k_last - 1;
    if (tid < 32) {
        const int64_t v = (DIR == MOT3D_DIR_OUT) ? ef_gallop<false, true>(p.tkey, r0, n, r0, kmin)
                                                 : ef_gallop<false, false>(p.tkey, 0, r1, r0, kmin);
        if (tid == 0) {
            s_range[0] = v;
            mbar_init(&s_bar, 1);
            fence_mbar_init();
        }
    } else if (tid < 64) {
        const int64_t v = (DIR == MOT3D_DIR_OUT) ? ef_gallop<true, true>(p.tkey, r0, n, r1, kmax)
                                                 : ef_gallop<true, false>(p.tkey, 0, r1, r1 - 1, kmax);
        if (tid == 32) s_range[1] = v;
    }
    for (int t = tid; t < EF_TABLE; t += EF_TPB) {
        s_seg_lo[t] = 0;
        s_seg_hi[t] = 0;
    }
    __syncthreads();
    const int64_t lo = s_range[0], hi = s_range[1] > s_range[0] ? s_range[1] : s_range[0];
    int m = (int)(hi - lo);
    if (hi - lo >= (1 << 24)) {  // candidate slots are packed into 24 bits
        if (tid == 0 && p.overflow) atomicOr(p.overflow, 2);
        m = (1 << 24) - 1;
    }
    const bool staged = m <= p.cap;
    const bool tabled = (kmax - kmin) < EF_TABLE;

    const int cap4 = p.cap + 4;
    double *b_x = (double *)ef_smem;
    double *b_y = b_x + cap4;
    double *b_z = b_y + cap4;  // only there when has_z (the launch sizes the buffer accordingly)
    double *b_rx = b_y + cap4 + (has_z ? cap4 : 0);  // the tile's own rows
    double *b_ry = b_rx + EF_TPB;
    double *b_rz = b_ry + EF_TPB;
    int32_t *b_perm = (int32_t *)(b_ry + EF_TPB + (has_z ? EF_TPB : 0));
    uint32_t *b_pk = (uint32_t *)(b_perm + cap4);
    uint32_t *cache = (staged && p.cache_j > 0) ? b_pk + p.acap : nullptr;
    const double *s_x = b_x, *s_y = b_y, *s_z = b_z;
    const int32_t *s_perm = b_perm;
    if (staged && m > 0) {
        uint32_t bulk = 0;
        s_x = b_x + stage_slice(p.spos, b_x, lo, m, &s_bar, &bulk);
        s_y = b_y + stage_slice(p.spos + n, b_y, lo, m, &s_bar, &bulk);
        if (has_z) s_z = b_z + stage_slice(p.spos + 2 * n, b_z, lo, m, &s_bar, &bulk);
        s_perm = b_perm + stage_slice(p.sperm, b_perm, lo, m, &s_bar, &bulk);
        if (tid == 0) mbar_expect_tx(&s_bar, bulk);
    }
    // frame segments of the candidate range, indexed by key - kmin (keys straight from global memory: coalesced, and
    // the TMA copies are in flight meanwhile)
    if (tabled) {
        for (int c = tid; c < m; c += EF_TPB) {
            const int k = p.tkey[lo + c];
            const int kp = c > 0 ? p.tkey[lo + c - 1] : k - 1;
            const int kn = c + 1 < m ? p.tkey[lo + c + 1] : k + 1;
            if (kp != k) s_seg_lo[k - kmin] = c;
            if (kn != k) s_seg_hi[k - kmin] = c + 1;
        }
    }
    // the row's own detection
    int key_r = 0;
    double xr = 0, yr = 0, zr = 0;
    if (active) {
        key_r = p.tkey[r];
        xr = p.pos[r];
        yr = p.pos[n + r];
        if (has_z) zr = p.pos[2 * n + r];
    }
    b_rx[tid] = xr;
    b_ry[tid] = yr;
    if (has_z) b_rz[tid] = zr;
    __syncthreads();  // table + ragged head/tail stores visible
    if (staged && m > 0) mbar_wait(&s_bar, 0);

    // ---- pass 1: count
    int cnt = 0;
    if (active) {
        if (staged)
            cnt = ef_row<DIR, false, true, false>(p, key_r, xr, yr, zr, (int)kmin, tabled, s_seg_lo, s_seg_hi, s_x, s_y, s_z,
                                                  s_perm, p.tkey + lo, m, cache, nullptr, nullptr, 0);
        else
            cnt = ef_row<DIR, false, false, false>(p, key_r, xr, yr, zr, (int)kmin, tabled, s_seg_lo, s_seg_hi, p.spos + lo,
                                                   p.spos + n + lo, p.spos + 2 * n + lo, p.sperm + lo, p.tkey + lo, m,
                                                   nullptr, nullptr, nullptr, 0);
    }
    int total;
    const int off = block_excl_scan(cnt, s_scan, &total);
    s_rowoff[tid] = off;
    if (tid == 0) s_rowoff[EF_TPB] = total;

    // ---- decoupled look-back over the tiles' arc counts: the aggregate is published right away, the walk over the
    //      predecessors (warp 0, 32 tiles per step) waits until the warp has done its share of pass 2 - by then the
    //      predecessors have usually published their prefix and one step is enough
    if (tid == 0) {
        if (tile == 0) {
            const long long b0 = p.arc_base_in ? (long long)*p.arc_base_in : 0;
            ef_st_status(p.status, EF_FLAG_PREFIX | (unsigned long long)(b0 + total));
            s_base = b0;
        } else {
            ef_st_status(p.status + tile, EF_FLAG_AGG | (unsigned long long)total);
        }
    }
    // ---- pass 2: the tile's arc list. Staged: 4-byte arcs (slot, jump, row) in shared memory. Else (candidates or
    //      arcs do not fit): built in place in col[] / cost[] at the tile's final offset (slot | jump, squared distance)
    const bool arcs_staged = staged && total <= p.acap;
    uint32_t *pk = b_pk;
    double *sv = nullptr;
    bool row_ok = true;
    if (!arcs_staged) {
        if (tid < 32 && tile > 0) ef_look_back(p, tile, total, &s_base);
        __syncthreads();
        const long long base = s_base;
        pk = (uint32_t *)(p.col + base);
        sv = (double *)(p.cost + base);
        row_ok = base + off + cnt <= p.edge_capacity;
        if (!row_ok && cnt > 0 && p.overflow) atomicOr(p.overflow, 1);
    }
    if (active && cnt > 0 && row_ok) {
        if (arcs_staged)
            ef_row<DIR, true, true, true>(p, key_r, xr, yr, zr, (int)kmin, tabled, s_seg_lo, s_seg_hi, s_x, s_y, s_z, s_perm,
                                          p.tkey + lo, m, cache, pk, nullptr, off);
        else if (staged)
            ef_row<DIR, true, true, false>(p, key_r, xr, yr, zr, (int)kmin, tabled, s_seg_lo, s_seg_hi, s_x, s_y, s_z, s_perm,
                                           p.tkey + lo, m, cache, pk, sv, off);
        else
            ef_row<DIR, true, false, false>(p, key_r, xr, yr, zr, (int)kmin, tabled, s_seg_lo, s_seg_hi, p.spos + lo,
                                            p.spos + n + lo, p.spos + 2 * n + lo, p.sperm + lo, p.tkey + lo, m, nullptr,
                                            pk, sv, off);
    }
    if (arcs_staged && tid < 32 && tile > 0) ef_look_back(p, tile, total, &s_base);
    __syncthreads();
    const long long base = s_base;
    if (active) p.row_ptr[r] = (int32_t)(base + off);
    if (r1 == p.row_end && tid == 0) {
        p.row_ptr[p.row_end] = (int32_t)(base + total);
        if (p.arc_total_out) *p.arc_total_out = base + total;
    }

    // ---- cost pass: one thread per arc, coalesced stores
    const int32_t *cperm = staged ? s_perm : p.sperm + lo;
    for (int i = tid; i < total; i += EF_TPB) {
        const long long e = base + i;
        if (e >= p.edge_capacity) {
            if (p.overflow) atomicOr(p.overflow, 1);
            break;
        }
        const uint32_t w32 = pk[i];
        int row, slot;
        uint32_t jm1;
        double s;
        if (arcs_staged) {
            slot = (int)(w32 & 0xffffu);
            jm1 = (w32 >> 16) & 0xffu;
            row = (int)(w32 >> 24);
            // the expression of the window test, so the same bits
            const double dx = b_rx[row] - s_x[slot];
            const double dy = b_ry[row] - s_y[slot];
            s = dx * dx;
            s = s + dy * dy;
            if (has_z) {
                const double dz = b_rz[row] - s_z[slot];
                s = s + dz * dz;
            }
        } else {
            // row of arc i: last row with offset <= i; rows of the in-place list that did not fit were never written
            row = 0;
            int hi_ = EF_TPB;
            while (hi_ - row > 1) {
                const int mid = (row + hi_) >> 1;
                if (s_rowoff[mid] <= i)
                    row = mid;
                else
                    hi_ = mid;
            }
            if (base + s_rowoff[row + 1] > p.edge_capacity) continue;
            slot = (int)(w32 & 0xffffffu);
            jm1 = w32 >> 24;
            s = sv[i];
        }
        const int64_t other = cperm[slot];
        double prod = distance_factor(s, p.sigma_distance_sq);
        double w;
        if (p.use_hist) {
            w = prod;  // mot3d_edge_appearance multiplies the histogram factor in first, as the reference does
        } else {
            if (p.use_bbox) {
                const int64_t rr = r0 + row;
                const int64_t a = (DIR == MOT3D_DIR_OUT) ? rr : other;
                const int64_t b = (DIR == MOT3D_DIR_OUT) ? other : rr;
                const double bss = 1.0 - bbox_similarity(p.bbox, n, a, b);
                prod = prod * exp(-(bss * bss) / p.sigma_bbox_sq);
            }
            w = p.jt.v[jm1] * prod;
            p.cost[e] = quantise_1e7(w);
        }
        p.col[e] = (int32_t)other;
        if (p.weight) p.weight[e] = w;
    }
}

int32_t check_dets_fused(const mot3d_dets *d, const mot3d_cost_spec *s) {
    M3_CHECK_ARG(d && s, "dets/spec is NULL");
    M3_CHECK_ARG(d->n >= 0 && d->n < ((int64_t)1 << 31), "detection count out of range");
    M3_CHECK_ARG(d->dim == 2 || d->dim == 3, "dim must be 2 or 3");
    M3_CHECK_ARG(s->max_jump >= 1 && s->max_jump <= MOT3D_MAX_JUMP, "max_jump must be in [1, MOT3D_MAX_JUMP]");
    M3_CHECK_ARG(d->n == 0 || (d->tkey && d->pos), "tkey/pos is NULL");
    M3_CHECK_ARG(d->n == 0 || (d->sorted_pos && d->sorted_perm),
                 "mot3d_edge_build needs the x-sorted copies (mot3d_frame_sort)");
    M3_CHECK_ARG(!s->use_bbox || d->bbox, "use_bbox set but bbox is NULL");
    return MOT3D_OK;
}

}  // namespace m3

using namespace m3;

extern "C" {

size_t mot3d_edge_build_workspace_bytes(int64_t n_rows) {
    if (n_rows < 0) return 0;
    const int64_t tiles = (n_rows + EF_TPB - 1) / EF_TPB;
    return (size_t)(tiles > 0 ? tiles : 1) * sizeof(unsigned long long) + 64;
}

int32_t mot3d_edge_build(const mot3d_dets *dets, const mot3d_cost_spec *spec, int32_t direction, int64_t row_begin,
                         int64_t row_end, int32_t *row_ptr_out, int64_t edge_capacity, int32_t *col_out,
                         int64_t *cost_out, double *weight_out, int32_t *overflow_flag, const int64_t *arc_base_in,
                         int64_t *arc_total_out, void *ws, size_t ws_bytes, void *stream) {
    int32_t rc = check_dets_fused(dets, spec);
    if (rc) return rc;
    M3_CHECK_ARG(direction == MOT3D_DIR_IN || direction == MOT3D_DIR_OUT, "bad direction");
    M3_CHECK_ARG(0 <= row_begin && row_begin <= row_end && row_end <= dets->n, "bad row range");
    M3_CHECK_ARG(row_ptr_out && edge_capacity >= 0, "row_ptr_out is NULL / negative capacity");
    M3_CHECK_ARG(edge_capacity == 0 || (col_out && cost_out), "col_out/cost_out is NULL");
    M3_CHECK_ARG(!(spec->use_hist || spec->defer_cost) || weight_out,
                 "use_hist / defer_cost need weight_out (finished by mot3d_edge_appearance[_views])");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t rows = row_end - row_begin;
    if (rows == 0) {
        // an empty range still closes the CSR where the previous launch left it
        if (arc_base_in) {
            M3_CUDA(cudaMemcpyAsync(row_ptr_out + row_end, arc_base_in, sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
            if (arc_total_out && arc_total_out != arc_base_in)
                M3_CUDA(cudaMemcpyAsync(arc_total_out, arc_base_in, sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
        } else {
            M3_CUDA(cudaMemsetAsync(row_ptr_out + row_end, 0, sizeof(int32_t), st));
            if (arc_total_out) M3_CUDA(cudaMemsetAsync(arc_total_out, 0, sizeof(int64_t), st));
        }
        return MOT3D_OK;
    }
    M3_CHECK_ARG(ws, "workspace is NULL");
    const size_t need = mot3d_edge_build_workspace_bytes(rows);
    if (ws_bytes < need) {
        set_error("edge build workspace too small: %zu < %zu bytes", ws_bytes, need);
        return MOT3D_E_CAPACITY;
    }
    const int64_t tiles = (rows + EF_TPB - 1) / EF_TPB;
    FusedParams p;
    memset(&p, 0, sizeof(p));
    p.n = dets->n;
    p.row_begin = row_begin;
    p.row_end = row_end;
    p.dim = dets->dim;
    p.max_jump = spec->max_jump;
    p.use_bbox = spec->use_bbox;
    p.use_hist = (spec->use_hist || spec->defer_cost) ? 1 : 0;  // = "distance factor only, an appearance pass finishes"
    p.tkey = dets->tkey;
    p.pos = dets->pos;
    p.spos = dets->sorted_pos;
    p.sperm = dets->sorted_perm;
    p.bbox = dets->bbox;
    p.sq_dist_max = spec->sq_dist_max;
    // a superset of the x-interval that can hold survivors: dx^2 <= s <= sq_dist_max
    const double w = spec->sq_dist_max > 0 ? sqrt(spec->sq_dist_max) : 0.0;
    p.x_window = w * (1.0 + 1e-12) + 1e-300;
    p.sigma_distance_sq = spec->sigma_distance_sq;
    p.sigma_bbox_sq = spec->sigma_bbox_sq;
    for (int i = 0; i < MOT3D_MAX_JUMP; i++) p.jt.v[i] = (i < spec->max_jump) ? spec->neg_exp_jump[i] : 0.0;
    p.row_ptr = row_ptr_out;
    p.edge_capacity = edge_capacity;
    p.col = col_out;
    p.cost = cost_out;
    p.weight = weight_out;
    p.overflow = overflow_flag;
    p.status = (unsigned long long *)ws;
    p.ticket = (int32_t *)((char *)ws + (size_t)tiles * sizeof(unsigned long long));
    p.arc_base_in = arc_base_in;
    p.arc_total_out = arc_total_out;
    // Shared memory per CTA, four CTAs per SM: candidates 20 (28 in 3-D) bytes each, the tile's rows 16 (24), arcs 4,
    // the pass-1 cache 4 bytes per row and frame. mot3d_set_option("edge_cap" / "edge_acap") overrides the split (tests).
    p.cache_j = spec->max_jump <= EF_CACHE_J ? spec->max_jump : 0;
    const int per_cand = 8 + 8 + (dets->dim == 3 ? 8 : 0) + 4;
    const int rows_b = EF_TPB * 8 * dets->dim;
    const int budget = 50 * 1024 - p.cache_j * EF_TPB * 4 - rows_b;
    int cap = 1280;
    if (dets->dim == 3) cap = 1024;
    if ((cap + 4) * per_cand > budget - 8192) cap = ((budget - 8192) / per_cand - 4) & ~3;  // room for 2048 arcs
    if (cap < 0) cap = 0;
    int acap = ((budget - (cap + 4) * per_cand - 16) / 4) & ~3;
    double ov;  // test hooks: force the staging modes (tests/test_gpu_parity.py)
    if (option(OPT_EDGE_CAP, &ov)) cap = (int)ov & ~3;
    if (option(OPT_EDGE_ACAP, &ov)) acap = (int)ov & ~3;
    if (cap > 65000) cap = 65000;  // slots of staged candidates are packed into 16 bits
    if (cap < 0) cap = 0;
    if (acap < 0) acap = 0;
    p.cap = cap;
    p.acap = acap;
    const size_t smem = (size_t)(cap + 4) * per_cand + (size_t)rows_b + (size_t)acap * 4 + (size_t)p.cache_j * EF_TPB * 4 + 16;
    M3_CUDA(cudaMemsetAsync(ws, 0, need, st));
    // (the shared-memory limit of a kernel is a per-device attribute: set when it has to grow, not on every call)
    static std::atomic<int> smem_set[MOT3D_MAX_DEVICES][2];
    int dev = 0;
    M3_CUDA(cudaGetDevice(&dev));
    const int di = direction == MOT3D_DIR_IN ? 0 : 1;
    const bool grow = dev < 0 || dev >= MOT3D_MAX_DEVICES || smem_set[dev][di].load(std::memory_order_relaxed) < (int)smem;
    if (direction == MOT3D_DIR_IN) {
        if (grow) M3_CUDA(cudaFuncSetAttribute(k_edge_build_fused<MOT3D_DIR_IN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (grow && dev >= 0 && dev < MOT3D_MAX_DEVICES) smem_set[dev][di].store((int)smem, std::memory_order_relaxed);
        k_edge_build_fused<MOT3D_DIR_IN><<<(unsigned)tiles, EF_TPB, smem, st>>>(p);
    } else {
        if (grow) M3_CUDA(cudaFuncSetAttribute(k_edge_build_fused<MOT3D_DIR_OUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (grow && dev >= 0 && dev < MOT3D_MAX_DEVICES) smem_set[dev][di].store((int)smem, std::memory_order_relaxed);
        k_edge_build_fused<MOT3D_DIR_OUT><<<(unsigned)tiles, EF_TPB, smem, st>>>(p);
    }
    M3_LAUNCH_CHECK("k_edge_build_fused");
    return MOT3D_OK;
}

}  // extern "C"
