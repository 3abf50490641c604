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
    // (4) order by (start frame, number): rank by counting
    for (int i = tid; i < Mk; i += TP_T) {
        const int fi = p.frame[p.track_det[p.w_keep[i]]];
        int rank = 0;
        for (int j = 0; j < Mk; j++) {
            const int fj = p.frame[p.track_det[p.w_keep[j]]];
            rank += (fj < fi || (fj == fi && j < i)) ? 1 : 0;
        }
        p.trk_first[rank] = p.w_keep[i];
        p.trk_len[rank] = p.w_len[i];
        p.key[rank] = fi;
    }
    __syncthreads();
    // (5) the windows at both ends: head = detections with index > last - window, tail = index < first + window
    //     (window 0 = None: the whole tracklet at both ends, mot3d/types.py:100-112); offsets by a scan, then the points
    int run_h = 0, run_t = 0;
    for (int base = 0; base < Mk; base += TP_T) {
        const int i = base + tid;
        int nh = 0, nt = 0;
        if (i < Mk) {
            const int first = p.trk_first[i], len = p.trk_len[i];
            if (p.window <= 0) {
                nh = nt = len;
            } else {
                const int f_last = p.frame[p.track_det[first + len - 1]], f_first = p.frame[p.track_det[first]];
                for (int k = 0; k < len; k++) {
                    const int f = p.frame[p.track_det[first + k]];
                    nh += f > f_last - p.window ? 1 : 0;
                    nt += f < f_first + p.window ? 1 : 0;
                }
            }
        }
        int tot_h, tot_t;
        const int eh = tp_block_excl_scan(nh, s_scan, &tot_h);
        const int et = tp_block_excl_scan(nt, s_scan, &tot_t);
        if (i < Mk) {
            p.head_ptr[i] = run_h + eh;
            p.tail_ptr[i] = run_t + et;
        }
        run_h += tot_h;
        run_t += tot_t;
    }
    if (tid == 0) {
        p.head_ptr[Mk] = run_h;
        p.tail_ptr[Mk] = run_t;
        p.n_tracklets[0] = Mk;
    }
    __syncthreads();
    for (int i = tid; i < Mk; i += TP_T) {
        const int first = p.trk_first[i], len = p.trk_len[i];
        const int f_last = p.frame[p.track_det[first + len - 1]], f_first = p.frame[p.track_det[first]];
        int h = p.head_ptr[i], t = p.tail_ptr[i];
        for (int k = 0; k < len; k++) {
            const int d = p.track_det[first + k];
            const int f = p.frame[d];
            if (p.window <= 0 || f > f_last - p.window) {
                p.head_idx[h] = f;
                for (int c = 0; c < p.dim; c++) p.head_pos[(size_t)c * p.n + h] = p.pos[(size_t)c * p.n + d];
                h++;
            }
            if (p.window <= 0 || f < f_first + p.window) {
                p.tail_idx[t] = f;
                for (int c = 0; c < p.dim; c++) p.tail_pos[(size_t)c * p.n + t] = p.pos[(size_t)c * p.n + d];
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
    M3_CHECK_ARG(n >= 0 && (dim == 2 || dim == 3), "n >= 0, dim 2 or 3");
    M3_CHECK_ARG(split_length >= 1, "split length must be >= 1");
    M3_CHECK_ARG(n_tracklets_out, "NULL output");
    if (n == 0) {
        M3_CUDA(cudaMemsetAsync(n_tracklets_out, 0, 4, (cudaStream_t)stream));
        return MOT3D_OK;
    }
    M3_CHECK_ARG(frame && pos && track_count && track_ptr && track_det, "NULL input");
    M3_CHECK_ARG(trk_first_out && trk_len_out && key_out && head_ptr_out && tail_ptr_out && head_idx_out && tail_idx_out &&
                     head_pos_out && tail_pos_out && ws,
                 "NULL output");
    if (ws_bytes < mot3d_tracklets_from_tracks_workspace_bytes(n)) {
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
    k_split_tracks<<<1, TP_T, 0, (cudaStream_t)stream>>>(p);
    M3_LAUNCH_CHECK("k_split_tracks");
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
