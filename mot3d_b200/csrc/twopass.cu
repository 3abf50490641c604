// Device-resident glue between the detection pass and the tracklet pass. Replaces, for one graph,
//   split_trajectory_modulo + remove_short_trajectories   (reference mot3d/utils/trajectory.py:216-248)
//   DetectionTracklet2D/3D(tracklet, window=...)           (mot3d/types.py:94-184: the head / tail windows)
//   concat_tracklets                                        (mot3d/utils/trajectory.py:11-12)
// The tracks of pass 1 (mot3d_extract_tracks) never leave the device: k_split_tracks cuts them where `index % length`
// wraps around, drops the pieces that are too short, orders the survivors by the frame they start in (a tracklet only
// links to tracklets that start after it ends, so that order is topological and the start frame is the solver's time
// key) and writes the mot3d_tracklets arrays the arc kernels of csrc/tracklet.cu read. k_out_to_in_csr turns their
// out-CSR into the in-CSR the solver pulls from; k_concat_tracklets turns the solver's chains of tracklets back into
// chains of detections. Single graph, single block each: these are a few thousand items.
#include "common.cuh"

namespace m3 {

constexpr int TP_T = 1024;

__device__ __forceinline__ int tp_block_excl_scan(int v, int *scratch, int *total) {
    return block_excl_scan(v, scratch, total);
}

struct SplitParams {
    int32_t n, dim;
    const int32_t *frame;
    const double *pos;
    const int32_t *track_count, *track_ptr, *track_det;
    int32_t length, th_length, window;
    int32_t *n_tracklets, *trk_first, *trk_len, *key;
    int32_t *head_ptr, *tail_ptr, *head_idx, *tail_idx;
    double *head_pos, *tail_pos;
    int32_t *w_flag, *w_first, *w_len, *w_keep;  // scratch [n + 1] each
    // trajectories carried over from earlier batches (the online loop: Scene.active) join the tracklet set as they are
    int32_t n_carry;
    const int32_t *carry_ptr, *carry_frame;  // [n_carry + 1], [points]
    const double *carry_pos;                 // [dim][carry_stride] planar
    int64_t carry_stride;
    int32_t *item;  // out [n_carry + kept]: what the tracklet at a rank is: c < n_carry = carried trajectory c, else
                    // n_carry + number of the new tracklet in track order (the order the reference lists them in)
    int64_t out_stride;  // planar stride of head_pos / tail_pos
};

__device__ __forceinline__ int py_mod(int a, int m) {
    const int r = a % m;
    return r < 0 ? r + m : r;
}

__global__ void __launch_bounds__(TP_T) k_split_tracks(const SplitParams p) {
    __shared__ int s_scan[33];
    const int tid = threadIdx.x;
    const int K = p.track_count[0];
    const int used = K > 0 ? p.track_ptr[K] : 0;  // detections on tracks, track after track
    // (1) where a tracklet starts: at a track's first detection, and wherever index % length falls
    for (int q = tid; q <= used; q += TP_T) p.w_flag[q] = 0;
    __syncthreads();
    for (int t = tid; t < K; t += TP_T) p.w_flag[p.track_ptr[t]] = 1;
    __syncthreads();
    for (int q = tid; q < used; q += TP_T)
        if (q > 0 && !p.w_flag[q] &&
            py_mod(p.frame[p.track_det[q]], p.length) < py_mod(p.frame[p.track_det[q - 1]], p.length))
            p.w_flag[q] = 1;
    __syncthreads();
    // (2) number them (first positions in w_first), then their lengths
    int M = 0;
    for (int base = 0; base < used; base += TP_T) {
        const int q = base + tid;
        const int f = q < used ? p.w_flag[q] : 0;
        int tot;
        const int ex_ = tp_block_excl_scan(f, s_scan, &tot);
        if (f) p.w_first[M + ex_] = q;
        M += tot;
    }
    __syncthreads();
    // (3) keep the ones with more than th_length detections; compact (still in track order)
    int Mk = 0;
    for (int base = 0; base < M; base += TP_T) {
        const int i = base + tid;
        int len = 0, first = 0;
        if (i < M) {
            first = p.w_first[i];
            len = (i + 1 < M ? p.w_first[i + 1] : used) - first;
        }
        const int f = (i < M && len > p.th_length) ? 1 : 0;
        int tot;
        const int ex_ = tp_block_excl_scan(f, s_scan, &tot);
        if (f) {
            p.w_keep[Mk + ex_] = first;
            p.w_len[Mk + ex_] = len;
        }
        Mk += tot;
    }
    __syncthreads();
    // (4) the tracklet set = carried trajectories, then the new tracklets (the reference's list order); ordered by
    //     (start frame, list index): rank by counting. For a carried trajectory first / len address carry_frame.
    const int A = p.n_carry, Mt = A + Mk;
    auto first_frame = [&](int it) {
        return it < A ? p.carry_frame[p.carry_ptr[it]] : p.frame[p.track_det[p.w_keep[it - A]]];
    };
    for (int i = tid; i < Mt; i += TP_T) {
        const int fi = first_frame(i);
        int rank = 0;
        for (int j = 0; j < Mt; j++) {
            const int fj = first_frame(j);
            rank += (fj < fi || (fj == fi && j < i)) ? 1 : 0;
        }
        p.trk_first[rank] = i < A ? p.carry_ptr[i] : p.w_keep[i - A];
        p.trk_len[rank] = i < A ? p.carry_ptr[i + 1] - p.carry_ptr[i] : p.w_len[i - A];
        p.key[rank] = fi;
        if (p.item) p.item[rank] = i;
    }
    __syncthreads();
    // (5) the windows at both ends: head = detections with index > last - window, tail = index < first + window
    //     (window 0 = None: the whole tracklet at both ends, mot3d/types.py:100-112); offsets by a scan, then the points
    auto is_carried = [&](int r) { return p.item && p.item[r] < A; };
    auto frame_of = [&](int r, int k) {
        return is_carried(r) ? p.carry_frame[p.trk_first[r] + k] : p.frame[p.track_det[p.trk_first[r] + k]];
    };
    int run_h = 0, run_t = 0;
    for (int base = 0; base < Mt; base += TP_T) {
        const int i = base + tid;
        int nh = 0, nt = 0;
        if (i < Mt) {
            const int len = p.trk_len[i];
            if (p.window <= 0) {
                nh = nt = len;
            } else {
                const int f_last = frame_of(i, len - 1), f_first = frame_of(i, 0);
                for (int k = 0; k < len; k++) {
                    const int f = frame_of(i, k);
                    nh += f > f_last - p.window ? 1 : 0;
                    nt += f < f_first + p.window ? 1 : 0;
                }
            }
        }
        int tot_h, tot_t;
        const int eh = tp_block_excl_scan(nh, s_scan, &tot_h);
        const int et = tp_block_excl_scan(nt, s_scan, &tot_t);
        if (i < Mt) {
            p.head_ptr[i] = run_h + eh;
            p.tail_ptr[i] = run_t + et;
        }
        run_h += tot_h;
        run_t += tot_t;
    }
    if (tid == 0) {
        p.head_ptr[Mt] = run_h;
        p.tail_ptr[Mt] = run_t;
        p.n_tracklets[0] = Mt;
    }
    __syncthreads();
    for (int i = tid; i < Mt; i += TP_T) {
        const int first = p.trk_first[i], len = p.trk_len[i];
        const bool carried = is_carried(i);
        const int f_last = frame_of(i, len - 1), f_first = frame_of(i, 0);
        int h = p.head_ptr[i], t = p.tail_ptr[i];
        for (int k = 0; k < len; k++) {
            const int d = carried ? first + k : p.track_det[first + k];
            const int f = carried ? p.carry_frame[d] : p.frame[d];
            const double *src = carried ? p.carry_pos : p.pos;
            const int64_t stride = carried ? p.carry_stride : (int64_t)p.n;
            if (p.window <= 0 || f > f_last - p.window) {
                p.head_idx[h] = f;
                for (int c = 0; c < p.dim; c++) p.head_pos[(size_t)c * p.out_stride + h] = src[(size_t)c * stride + d];
                h++;
            }
            if (p.window <= 0 || f < f_first + p.window) {
                p.tail_idx[t] = f;
                for (int c = 0; c < p.dim; c++) p.tail_pos[(size_t)c * p.out_stride + t] = src[(size_t)c * stride + d];
                t++;
            }
        }
    }
}

// out-CSR (rows = arc tails) -> in-CSR (rows = arc heads) of one small graph, and its per-node costs. The order of the
// sources inside a row does not matter to the solver (ties go to the lowest source slot whatever the order).
__global__ void __launch_bounds__(TP_T) k_out_to_in_csr(const int32_t *__restrict__ n_nodes_dev, const int32_t *__restrict__ row_ptr,
                                                        const int32_t *__restrict__ col, const int64_t *__restrict__ cost,
                                                        int64_t entry, int64_t node, int64_t exit_,
                                                        int32_t *__restrict__ in_ptr, int32_t *__restrict__ in_src,
                                                        int64_t *__restrict__ in_cost, int64_t *__restrict__ en,
                                                        int64_t *__restrict__ cn, int64_t *__restrict__ ex,
                                                        int32_t *__restrict__ graph_ptr, int32_t *__restrict__ cursor) {
    __shared__ int s_scan[33];
    const int tid = threadIdx.x;
    const int M = n_nodes_dev[0];
    const int E = M > 0 ? row_ptr[M] : 0;
    for (int i = tid; i <= M; i += TP_T) cursor[i] = 0;
    for (int i = tid; i < M; i += TP_T) {
        en[i] = entry;
        cn[i] = node;
        ex[i] = exit_;
    }
    if (tid == 0) {
        graph_ptr[0] = 0;
        graph_ptr[1] = M;
    }
    __syncthreads();
    for (int e = tid; e < E; e += TP_T) atomicAdd(&cursor[col[e]], 1);
    __syncthreads();
    int run = 0;
    for (int base = 0; base < M; base += TP_T) {
        const int i = base + tid;
        const int v = i < M ? cursor[i] : 0;
        int tot;
        const int ex_ = tp_block_excl_scan(v, s_scan, &tot);
        if (i < M) {
            in_ptr[i] = run + ex_;
            cursor[i] = run + ex_;
        }
        run += tot;
    }
    if (tid == 0) in_ptr[M] = run;
    __syncthreads();
    for (int r = tid; r < M; r += TP_T)
        for (int e = row_ptr[r]; e < row_ptr[r + 1]; e++) {
            const int at = atomicAdd(&cursor[col[e]], 1);
            in_src[at] = r;
            in_cost[at] = cost[e];
        }
}

// chains of tracklets -> chains of detections (tracklets of a chain follow each other in time)
__global__ void __launch_bounds__(TP_T) k_concat_tracklets(const int32_t *__restrict__ track_count,
                                                           const int32_t *__restrict__ track_ptr,
                                                           const int32_t *__restrict__ track_trk,
                                                           const int32_t *__restrict__ trk_first,
                                                           const int32_t *__restrict__ trk_len,
                                                           const int32_t *__restrict__ det_of, int32_t *__restrict__ out_count,
                                                           int32_t *__restrict__ out_ptr, int32_t *__restrict__ out_det) {
    __shared__ int s_scan[33];
    const int tid = threadIdx.x;
    const int K = track_count[0];
    int run = 0;
    for (int base = 0; base < K; base += TP_T) {
        const int t = base + tid;
        int len = 0;
        if (t < K)
            for (int k = track_ptr[t]; k < track_ptr[t + 1]; k++) len += trk_len[track_trk[k]];
        int tot;
        const int ex_ = tp_block_excl_scan(len, s_scan, &tot);
        if (t < K) {
            int at = run + ex_;
            out_ptr[t] = at;
            for (int k = track_ptr[t]; k < track_ptr[t + 1]; k++) {
                const int i = track_trk[k];
                for (int q = 0; q < trk_len[i]; q++) out_det[at++] = det_of[trk_first[i] + q];
            }
        }
        run += tot;
    }
    if (tid == 0) {
        out_ptr[K] = run;
        out_count[0] = K;
    }
}

// The online loop's write-back: every output trajectory is a chain of tracklets (carried
// trajectories and pieces of this batch's tracks); its points are the points of the chain in time order with the
// skipped frames filled by linear interpolation of position and box -- concat_tracklets + interpolate_trajectory
// (reference mot3d/utils/trajectory.py:11-69: v1 + (v2 - v1) / (t2 - t1) * (t - t1)). One block per trajectory, one
// thread per source point: it writes its own point and the fill up to the next one.
struct OnlineConcatParams {
    int32_t n_out;
    const int32_t *chain_ptr, *chain;  // [n_out + 1], ranks of the tracklets of every chain
    const int32_t *out_ptr;            // [n_out + 1] first point of every output trajectory
    const int32_t *item, *trk_first, *trk_len;
    int32_t n_carry;
    // this batch
    int32_t n;
    const int32_t *track_det, *frame;
    const double *pos, *bbox;  // planar, stride n
    // carried
    const int32_t *carry_frame;
    const double *carry_pos, *carry_bbox;  // planar, stride carry_stride
    int64_t carry_stride;
    int32_t dim;
    // out
    int32_t *new_frame;
    double *new_pos, *new_bbox;  // planar, stride new_stride
    int64_t new_stride;
    int32_t *error;  // |= 1 when a chain does not fill its range exactly
    // colour histograms of the points (optional): hist_len doubles per point; filled points have none
    int32_t hist_len;
    const double *hist;        // this batch [n][hist_len] or NULL (no detection has one)
    const double *carry_hist;  // [carried points][hist_len]
    const uint8_t *carry_has;
    double *new_hist;
    uint8_t *new_has;
};

constexpr int OC_T = 128;
constexpr int OC_CHAIN = 256;  // chain prefix kept in shared memory; longer chains re-walk the list

__global__ void __launch_bounds__(OC_T) k_online_concat(const OnlineConcatParams p) {
    __shared__ int s_pre[OC_CHAIN + 1];
    const int t = blockIdx.x, tid = threadIdx.x;
    const int c0 = p.chain_ptr[t], c1 = p.chain_ptr[t + 1], nc = c1 - c0;
    if (tid == 0) {
        int run = 0;
        for (int k = 0; k < nc; k++) {
            if (k <= OC_CHAIN) s_pre[k] = run;
            run += p.trk_len[p.chain[c0 + k]];
        }
        if (nc <= OC_CHAIN) s_pre[nc] = run;
        else s_pre[OC_CHAIN] = run;  // total, whatever the length
    }
    __syncthreads();
    const int total = s_pre[nc <= OC_CHAIN ? nc : OC_CHAIN];
    const int o0 = p.out_ptr[t], o1 = p.out_ptr[t + 1];
    // point k of the chain -> (source, index there)
    auto locate = [&](int k, bool *carried) {
        int j = 0, base = 0;
        if (nc <= OC_CHAIN) {
            while (j + 1 < nc && s_pre[j + 1] <= k) j++;
            base = s_pre[j];
        } else {
            while (j + 1 < nc && base + p.trk_len[p.chain[c0 + j]] <= k) base += p.trk_len[p.chain[c0 + j++]];
        }
        const int r = p.chain[c0 + j];
        *carried = p.item[r] < p.n_carry;
        const int at = p.trk_first[r] + (k - base);
        return *carried ? at : p.track_det[at];
    };
    auto load = [&](int d, bool carried, int *f, double *v) {
        if (carried) {
            *f = p.carry_frame[d];
            for (int c = 0; c < p.dim; c++) v[c] = p.carry_pos[(size_t)c * p.carry_stride + d];
            for (int c = 0; c < 4; c++) v[3 + c] = p.carry_bbox[(size_t)c * p.carry_stride + d];
        } else {
            *f = p.frame[d];
            for (int c = 0; c < p.dim; c++) v[c] = p.pos[(size_t)c * p.n + d];
            for (int c = 0; c < 4; c++) v[3 + c] = p.bbox[(size_t)c * p.n + d];
        }
    };
    int f_first = 0;
    {
        bool cr;
        double v[7];
        const int d = locate(0, &cr);
        load(d, cr, &f_first, v);
    }
    for (int k = tid; k < total; k += OC_T) {
        bool cr1, cr2;
        int f1, f2;
        double v1[7], v2[7];
        const int d1 = locate(k, &cr1);
        load(d1, cr1, &f1, v1);
        int at = o0 + (f1 - f_first);
        if (at < o0 || at >= o1) {
            atomicOr(p.error, 1);
            continue;
        }
        p.new_frame[at] = f1;
        for (int c = 0; c < p.dim; c++) p.new_pos[(size_t)c * p.new_stride + at] = v1[c];
        for (int c = 0; c < 4; c++) p.new_bbox[(size_t)c * p.new_stride + at] = v1[3 + c];
        if (p.new_has) {
            const bool has = cr1 ? (p.carry_has[d1] != 0) : (p.hist != nullptr);
            p.new_has[at] = has ? 1 : 0;
            if (has) {
                const double *src = cr1 ? p.carry_hist + (size_t)d1 * p.hist_len : p.hist + (size_t)d1 * p.hist_len;
                for (int c = 0; c < p.hist_len; c++) p.new_hist[(size_t)at * p.hist_len + c] = src[c];
            }
        }
        if (k + 1 < total) {
            const int d2 = locate(k + 1, &cr2);
            load(d2, cr2, &f2, v2);
            if (f2 <= f1 || at + (f2 - f1) >= o1) {
                atomicOr(p.error, 1);
                continue;
            }
            const double dt = (double)(f2 - f1);
            for (int f = f1 + 1; f < f2; f++) {
                const int a2 = at + (f - f1);
                const double x = (double)(f - f1);
                p.new_frame[a2] = f;
                if (p.new_has) p.new_has[a2] = 0;
                for (int c = 0; c < p.dim; c++)
                    p.new_pos[(size_t)c * p.new_stride + a2] = (v2[c] - v1[c]) / dt * x + v1[c];
                for (int c = 0; c < 4; c++)
                    p.new_bbox[(size_t)c * p.new_stride + a2] = (v2[3 + c] - v1[3 + c]) / dt * x + v1[3 + c];
            }
        } else if (at != o1 - 1) {
            atomicOr(p.error, 1);
        }
    }
}

// Median colour histogram of the head and of the tail window of every tracklet: what DetectionTracklet2D keeps as
// head.color_histogram / tail.color_histogram (reference mot3d/types.py:69-86,109-125: np.median over the detections of
// the window that have a histogram; none of them -> None). One block per (tracklet, end), one thread per bin; the
// median by counting ranks (windows are a handful of points; a long carried trajectory with window None still works).
struct WindowHistParams {
    const int32_t *item, *trk_first, *trk_len;
    int32_t n_carry, window, hist_len;
    const int32_t *track_det, *frame;
    const double *hist;
    const int32_t *carry_frame;
    const double *carry_hist;
    const uint8_t *carry_has;
    double *head_hist, *tail_hist;
    uint8_t *head_has, *tail_has;
};

__global__ void __launch_bounds__(128) k_window_median_hist(const WindowHistParams p) {
    const int r = blockIdx.x >> 1, tail = blockIdx.x & 1;
    const int first = p.trk_first[r], len = p.trk_len[r];
    const bool carried = p.item && p.item[r] < p.n_carry;
    auto frame_of = [&](int k) { return carried ? p.carry_frame[first + k] : p.frame[p.track_det[first + k]]; };
    // the window is a run of points at one end (frames ascend along a tracklet)
    int lo = 0, hi = len;
    if (p.window > 0) {
        if (tail) {
            const int cut = frame_of(0) + p.window;
            hi = 0;
            while (hi < len && frame_of(hi) < cut) hi++;
        } else {
            const int cut = frame_of(len - 1) - p.window;
            lo = len;
            while (lo > 0 && frame_of(lo - 1) > cut) lo--;
        }
    }
    auto has = [&](int k) { return carried ? p.carry_has[first + k] != 0 : p.hist != nullptr; };
    auto row = [&](int k) {
        return carried ? p.carry_hist + (size_t)(first + k) * p.hist_len
                       : p.hist + (size_t)p.track_det[first + k] * p.hist_len;
    };
    int m = 0;
    for (int k = lo; k < hi; k++) m += has(k) ? 1 : 0;
    double *out = (tail ? p.tail_hist : p.head_hist) + (size_t)r * p.hist_len;
    if (threadIdx.x == 0) (tail ? p.tail_has : p.head_has)[r] = m > 0 ? 1 : 0;
    if (m == 0) return;
    const int want_a = (m - 1) >> 1, want_b = m >> 1;  // the middle one(s) in sorted order
    for (int c = threadIdx.x; c < p.hist_len; c += blockDim.x) {
        double va = 0.0, vb = 0.0;
        for (int i = lo; i < hi; i++) {
            if (!has(i)) continue;
            const double vi = row(i)[c];
            int rank = 0;
            for (int j = lo; j < hi; j++) {
                if (!has(j)) continue;
                const double vj = row(j)[c];
                rank += (vj < vi || (vj == vi && j < i)) ? 1 : 0;
            }
            if (rank == want_a) va = vi;
            if (rank == want_b) vb = vi;
        }
        out[c] = want_a == want_b ? va : (va + vb) / 2.0;
    }
}

// first / last frame of the tracklet at every rank (the host lays the output trajectories out from these)
__global__ void k_tracklet_spans(const int32_t *__restrict__ n_tracklets, const int32_t *__restrict__ item,
                                 const int32_t *__restrict__ trk_first, const int32_t *__restrict__ trk_len,
                                 int32_t n_carry, const int32_t *__restrict__ carry_frame,
                                 const int32_t *__restrict__ track_det, const int32_t *__restrict__ frame,
                                 int32_t *__restrict__ span) {
    const int M = n_tracklets[0];
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < M; r += gridDim.x * blockDim.x) {
        const int first = trk_first[r], last = first + trk_len[r] - 1;
        const bool carried = item[r] < n_carry;
        span[2 * r] = carried ? carry_frame[first] : frame[track_det[first]];
        span[2 * r + 1] = carried ? carry_frame[last] : frame[track_det[last]];
    }
}

}  // namespace m3

using namespace m3;

extern "C" {

size_t mot3d_tracklets_from_tracks_workspace_bytes(int64_t n) {
    if (n < 0) return 0;
    return (size_t)(n + 1) * 4 * 4 + 256;
}

int32_t mot3d_tracklets_from_tracks(int32_t n, int32_t dim, const int32_t *frame, const double *pos,
                                    const int32_t *track_count, const int32_t *track_ptr, const int32_t *track_det,
                                    int32_t split_length, int32_t th_length, int32_t window, int32_t *n_tracklets_out,
                                    int32_t *trk_first_out, int32_t *trk_len_out, int32_t *key_out, int32_t *head_ptr_out,
                                    int32_t *tail_ptr_out, int32_t *head_idx_out, int32_t *tail_idx_out,
                                    double *head_pos_out, double *tail_pos_out, void *ws, size_t ws_bytes, void *stream) {
    if (n == 0) {
        M3_CHECK_ARG(n_tracklets_out, "NULL output");
        M3_CUDA(cudaMemsetAsync(n_tracklets_out, 0, 4, (cudaStream_t)stream));
        return MOT3D_OK;
    }
    return mot3d_online_tracklets(n, dim, frame, pos, track_count, track_ptr, track_det, split_length, th_length, window,
                                  0, nullptr, nullptr, nullptr, 0, n, n_tracklets_out, nullptr, trk_first_out,
                                  trk_len_out, key_out, head_ptr_out, tail_ptr_out, head_idx_out, tail_idx_out,
                                  head_pos_out, tail_pos_out, nullptr, ws, ws_bytes, stream);
}

int32_t mot3d_online_tracklets(int32_t n, int32_t dim, const int32_t *frame, const double *pos,
                               const int32_t *track_count, const int32_t *track_ptr, const int32_t *track_det,
                               int32_t split_length, int32_t th_length, int32_t window, int32_t n_carry,
                               const int32_t *carry_ptr, const int32_t *carry_frame, const double *carry_pos,
                               int64_t carry_stride, int64_t out_stride, int32_t *n_tracklets_out, int32_t *item_out,
                               int32_t *trk_first_out, int32_t *trk_len_out, int32_t *key_out, int32_t *head_ptr_out,
                               int32_t *tail_ptr_out, int32_t *head_idx_out, int32_t *tail_idx_out, double *head_pos_out,
                               double *tail_pos_out, int32_t *span_out, void *ws, size_t ws_bytes, void *stream) {
    M3_CHECK_ARG(n >= 0 && (dim == 2 || dim == 3), "n >= 0, dim 2 or 3");
    M3_CHECK_ARG(split_length >= 1, "split length must be >= 1");
    M3_CHECK_ARG(n_tracklets_out, "NULL output");
    M3_CHECK_ARG(n_carry >= 0, "negative number of carried trajectories");
    M3_CHECK_ARG(n_carry == 0 || (carry_ptr && carry_frame && carry_pos && item_out), "NULL carried input");
    M3_CHECK_ARG(out_stride >= 1, "out_stride");
    M3_CHECK_ARG(track_count && track_ptr, "NULL input");
    M3_CHECK_ARG(n == 0 || (frame && pos && track_det), "NULL input");
    M3_CHECK_ARG(trk_first_out && trk_len_out && key_out && head_ptr_out && tail_ptr_out && head_idx_out && tail_idx_out &&
                     head_pos_out && tail_pos_out && ws,
                 "NULL output");
    if (!ws || ws_bytes < mot3d_tracklets_from_tracks_workspace_bytes(n)) {
        set_error("tracklet workspace too small");
        return MOT3D_E_CAPACITY;
    }
    SplitParams p;
    p.n = n;
    p.dim = dim;
    p.frame = frame;
    p.pos = pos;
    p.track_count = track_count;
    p.track_ptr = track_ptr;
    p.track_det = track_det;
    p.length = split_length;
    p.th_length = th_length;
    p.window = window;
    p.n_tracklets = n_tracklets_out;
    p.trk_first = trk_first_out;
    p.trk_len = trk_len_out;
    p.key = key_out;
    p.head_ptr = head_ptr_out;
    p.tail_ptr = tail_ptr_out;
    p.head_idx = head_idx_out;
    p.tail_idx = tail_idx_out;
    p.head_pos = head_pos_out;
    p.tail_pos = tail_pos_out;
    int32_t *w = (int32_t *)ws;
    p.w_flag = w;
    p.w_first = w + (n + 1);
    p.w_len = w + 2 * (size_t)(n + 1);
    p.w_keep = w + 3 * (size_t)(n + 1);
    p.n_carry = n_carry;
    p.carry_ptr = carry_ptr;
    p.carry_frame = carry_frame;
    p.carry_pos = carry_pos;
    p.carry_stride = carry_stride;
    p.item = item_out;
    p.out_stride = out_stride;
    k_split_tracks<<<1, TP_T, 0, (cudaStream_t)stream>>>(p);
    M3_LAUNCH_CHECK("k_split_tracks");
    if (span_out) {
        M3_CHECK_ARG(item_out, "spans need the item array");
        k_tracklet_spans<<<4, 256, 0, (cudaStream_t)stream>>>(n_tracklets_out, item_out, trk_first_out, trk_len_out, n_carry,
                                                             carry_frame, track_det, frame, span_out);
        M3_LAUNCH_CHECK("k_tracklet_spans");
    }
    return MOT3D_OK;
}

int32_t mot3d_online_concat(int32_t n_out, const int32_t *chain_ptr, const int32_t *chain, const int32_t *out_ptr,
                            const int32_t *item, const int32_t *trk_first, const int32_t *trk_len, int32_t n_carry,
                            int32_t n, int32_t dim, const int32_t *track_det, const int32_t *frame, const double *pos,
                            const double *bbox, const int32_t *carry_frame, const double *carry_pos,
                            const double *carry_bbox, int64_t carry_stride, int32_t *new_frame, double *new_pos,
                            double *new_bbox, int64_t new_stride, int32_t hist_len, const double *hist,
                            const double *carry_hist, const uint8_t *carry_has, double *new_hist, uint8_t *new_has,
                            int32_t *error_flag, void *stream) {
    M3_CHECK_ARG(n_out >= 0 && (dim == 2 || dim == 3), "n_out >= 0, dim 2 or 3");
    M3_CHECK_ARG(hist_len >= 0 && (hist_len == 0 || (new_hist && new_has)), "histograms: NULL output");
    M3_CHECK_ARG(hist_len == 0 || n_carry == 0 || (carry_hist && carry_has), "histograms: NULL carried input");
    if (n_out == 0) return MOT3D_OK;
    M3_CHECK_ARG(chain_ptr && chain && out_ptr && item && trk_first && trk_len && new_frame && new_pos && new_bbox &&
                     error_flag,
                 "NULL pointer");
    M3_CHECK_ARG(n_carry == 0 || (carry_frame && carry_pos && carry_bbox), "NULL carried input");
    M3_CHECK_ARG(n == 0 || (track_det && frame && pos && bbox), "NULL batch input");
    OnlineConcatParams p;
    p.n_out = n_out;
    p.chain_ptr = chain_ptr;
    p.chain = chain;
    p.out_ptr = out_ptr;
    p.item = item;
    p.trk_first = trk_first;
    p.trk_len = trk_len;
    p.n_carry = n_carry;
    p.n = n;
    p.track_det = track_det;
    p.frame = frame;
    p.pos = pos;
    p.bbox = bbox;
    p.carry_frame = carry_frame;
    p.carry_pos = carry_pos;
    p.carry_bbox = carry_bbox;
    p.carry_stride = carry_stride;
    p.dim = dim;
    p.new_frame = new_frame;
    p.new_pos = new_pos;
    p.new_bbox = new_bbox;
    p.new_stride = new_stride;
    p.error = error_flag;
    p.hist_len = hist_len;
    p.hist = hist;
    p.carry_hist = carry_hist;
    p.carry_has = carry_has;
    p.new_hist = hist_len ? new_hist : nullptr;
    p.new_has = hist_len ? new_has : nullptr;
    k_online_concat<<<n_out, OC_T, 0, (cudaStream_t)stream>>>(p);
    M3_LAUNCH_CHECK("k_online_concat");
    return MOT3D_OK;
}

int32_t mot3d_tracklet_window_hist(int32_t n_tracklets, const int32_t *item, const int32_t *trk_first,
                                   const int32_t *trk_len, int32_t n_carry, int32_t window, int32_t hist_len,
                                   const int32_t *track_det, const int32_t *frame, const double *hist,
                                   const int32_t *carry_frame, const double *carry_hist, const uint8_t *carry_has,
                                   double *head_hist_out, double *tail_hist_out, uint8_t *head_has_out,
                                   uint8_t *tail_has_out, void *stream) {
    M3_CHECK_ARG(n_tracklets >= 0 && hist_len >= 1 && n_carry >= 0, "sizes");
    if (n_tracklets == 0) return MOT3D_OK;
    M3_CHECK_ARG(trk_first && trk_len && head_hist_out && tail_hist_out && head_has_out && tail_has_out, "NULL pointer");
    M3_CHECK_ARG(n_carry == 0 || (item && carry_frame && carry_hist && carry_has), "NULL carried input");
    WindowHistParams p;
    p.item = item;
    p.trk_first = trk_first;
    p.trk_len = trk_len;
    p.n_carry = n_carry;
    p.window = window;
    p.hist_len = hist_len;
    p.track_det = track_det;
    p.frame = frame;
    p.hist = hist;
    p.carry_frame = carry_frame;
    p.carry_hist = carry_hist;
    p.carry_has = carry_has;
    p.head_hist = head_hist_out;
    p.tail_hist = tail_hist_out;
    p.head_has = head_has_out;
    p.tail_has = tail_has_out;
    k_window_median_hist<<<2 * n_tracklets, 128, 0, (cudaStream_t)stream>>>(p);
    M3_LAUNCH_CHECK("k_window_median_hist");
    return MOT3D_OK;
}

int32_t mot3d_tracklet_graph_in_csr(const int32_t *n_nodes_dev, int32_t n_nodes_max, const int32_t *row_ptr,
                                    const int32_t *col, const int64_t *cost, int64_t entry_cost, int64_t node_cost,
                                    int64_t exit_cost, int32_t *in_ptr_out, int32_t *in_src_out, int64_t *in_cost_out,
                                    int64_t *entry_out, int64_t *node_out, int64_t *exit_out, int32_t *graph_ptr_out,
                                    int32_t *ws, void *stream) {
    M3_CHECK_ARG(n_nodes_max >= 0, "negative size");
    M3_CHECK_ARG(n_nodes_dev && row_ptr && in_ptr_out && entry_out && node_out && exit_out && graph_ptr_out && ws,
                 "NULL pointer");
    k_out_to_in_csr<<<1, TP_T, 0, (cudaStream_t)stream>>>(n_nodes_dev, row_ptr, col, cost, entry_cost, node_cost,
                                                         exit_cost, in_ptr_out, in_src_out, in_cost_out, entry_out,
                                                         node_out, exit_out, graph_ptr_out, ws);
    M3_LAUNCH_CHECK("k_out_to_in_csr");
    return MOT3D_OK;
}

int32_t mot3d_concat_tracklets(const int32_t *track_count, const int32_t *track_ptr, const int32_t *track_trk,
                               const int32_t *trk_first, const int32_t *trk_len, const int32_t *track_det_pass1,
                               int32_t *out_count, int32_t *out_ptr, int32_t *out_det, void *stream) {
    M3_CHECK_ARG(track_count && track_ptr && track_trk && trk_first && trk_len && track_det_pass1 && out_count &&
                     out_ptr && out_det,
                 "NULL pointer");
    k_concat_tracklets<<<1, TP_T, 0, (cudaStream_t)stream>>>(track_count, track_ptr, track_trk, trk_first, trk_len,
                                                            track_det_pass1, out_count, out_ptr, out_det);
    M3_LAUNCH_CHECK("k_concat_tracklets");
    return MOT3D_OK;
}

}  // extern "C"
