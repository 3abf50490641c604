// Track gather: the only communication of the path (whole graphs are sharded over the GPUs of a box, no
// data-path collective; the ragged track outputs go to rank 0).
//
//   payload     one contiguous int32 buffer of FIXED size per rank, so that no size has to travel to the host before
//               the copy is issued:  [ K_g | L_g | track_det ]  = 2 G + n words, where K_g = tracks of graph g, L_g =
//               detections on them, and the first detection of every track carries MOT3D_TRACK_START_BIT. The
//               offsets of mot3d_extract_tracks (track_ptr) are redundant with those bits and are rebuilt by
//               mot3d_unpack_tracks on the receiving side.
//   transport   rank 0 owns one receive slot per rank in a buffer allocated by this library with cudaMalloc and
//               exported through CUDA IPC; every other rank opens it and PUTS its payload with cudaMemcpyAsync
//               (device-to-device over NVLink, executed by the copy engines: no SM is used, so the transfer neither
//               slows the solver down nor is starved by its persistent blocks), followed by a 4-byte put of the step
//               number into the slot's flag word. Rank 0 waits for the flags with a one-thread kernel on its stream.
//
// The reference has no counterpart (it is a single-process CPU library); the payload replaces the Python lists
// solve_graph returns (mot3d/mot3d.py:208-211) as the interchange format between ranks.
#include <string.h>

#include "common.cuh"

namespace m3 {

// one block per graph: counts, lengths, detections with the start bits
__global__ void __launch_bounds__(256) k_pack_tracks(int n_graphs, const int32_t *__restrict__ graph_ptr,
                                                     const int32_t *__restrict__ track_count,
                                                     const int32_t *__restrict__ track_ptr,
                                                     const int32_t *__restrict__ track_det, int32_t *__restrict__ out) {
    const int g = blockIdx.x;
    const int g0 = graph_ptr[g], n_total = graph_ptr[n_graphs];
    const int K = track_count[g];
    const int32_t *tp = track_ptr + g0 + g;
    const int L = K > 0 ? tp[K] : 0;
    int32_t *det = out + 2 * (size_t)n_graphs + g0;
    if (threadIdx.x == 0) {
        out[g] = K;
        out[n_graphs + g] = L;
    }
    for (int i = threadIdx.x; i < L; i += blockDim.x) det[i] = track_det[g0 + i];
    // (entries beyond L are not part of the payload's meaning; they are zeroed so that the buffer is reproducible)
    const int n = (g + 1 < n_graphs ? graph_ptr[g + 1] : n_total) - g0;
    for (int i = L + threadIdx.x; i < n; i += blockDim.x) det[i] = 0;
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += blockDim.x) det[tp[k]] |= (int32_t)MOT3D_TRACK_START_BIT;
}

// one block per graph: track_ptr from the start bits (block scan over the graph's L entries), bits cleared in track_det
__global__ void __launch_bounds__(256) k_unpack_tracks(int n_graphs, const int32_t *__restrict__ graph_ptr,
                                                       const int32_t *__restrict__ payload,
                                                       int32_t *__restrict__ track_count, int32_t *__restrict__ track_ptr,
                                                       int32_t *__restrict__ track_det) {
    __shared__ int s_scan[33];
    const int g = blockIdx.x;
    const int g0 = graph_ptr[g];
    const int K = payload[g], L = payload[n_graphs + g];
    const int32_t *det = payload + 2 * (size_t)n_graphs + g0;
    int32_t *tp = track_ptr + g0 + g;
    int run = 0;
    for (int base = 0; base < L; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const int v = i < L ? det[i] : 0;
        const int f = (i < L && (v & (int32_t)MOT3D_TRACK_START_BIT)) ? 1 : 0;
        int tot;
        const int ex = block_excl_scan(f, s_scan, &tot);
        if (f) tp[run + ex] = i;
        if (i < L) track_det[g0 + i] = v & ~(int32_t)MOT3D_TRACK_START_BIT;
        run += tot;
    }
    if (threadIdx.x == 0) {
        tp[K] = L;
        track_count[g] = K;
    }
}

__global__ void k_set_word(int32_t *dst, int32_t v) {
    *dst = v;
    __threadfence_system();
}

// a sender: spin until the owner's counter (peer memory, read over NVLink) has reached `need`
__global__ void k_wait_word(const int32_t *word, int32_t need, int32_t *timeout) {
    const long long t0 = clock64();
    for (;;) {
        int32_t v;
        asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(word) : "memory");
        if (v - need >= 0) return;
        if (clock64() - t0 > 4000000000ll) {
            *timeout = 1;
            return;
        }
        __nanosleep(500);
    }
}

// rank 0: spin until every rank's flag has reached `step` (the puts of all ranks have landed); gives up after ~2 s and
// raises *timeout so that a dead peer cannot hang the stream for ever
__global__ void k_wait_flags(const int32_t *flags, int stride_words, int world, int32_t step, int32_t *timeout) {
    const long long t0 = clock64();
    for (int r = 1; r < world; r++) {
        const int32_t *f = flags + (size_t)r * stride_words;
        for (;;) {
            int32_t v;
            asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
            if (v - step >= 0) break;
            if (clock64() - t0 > 4000000000ll) {
                *timeout = r;
                return;
            }
            __nanosleep(200);
        }
    }
}

}  // namespace m3

using namespace m3;

extern "C" {

int64_t mot3d_track_payload_words(int64_t n_total, int32_t n_graphs) {
    if (n_total < 0 || n_graphs < 0) return 0;
    return 2 * (int64_t)n_graphs + n_total;
}

int32_t mot3d_pack_tracks(int32_t n_graphs, const int32_t *graph_ptr, const int32_t *track_count,
                          const int32_t *track_ptr, const int32_t *track_det, int32_t *payload_out, void *stream) {
    M3_CHECK_ARG(n_graphs >= 0, "n_graphs < 0");
    if (n_graphs == 0) return MOT3D_OK;
    M3_CHECK_ARG(graph_ptr && track_count && track_ptr && track_det && payload_out, "NULL pointer");
    k_pack_tracks<<<n_graphs, 256, 0, (cudaStream_t)stream>>>(n_graphs, graph_ptr, track_count, track_ptr, track_det,
                                                             payload_out);
    M3_LAUNCH_CHECK("k_pack_tracks");
    return MOT3D_OK;
}

int32_t mot3d_unpack_tracks(int32_t n_graphs, const int32_t *graph_ptr, const int32_t *payload,
                            int32_t *track_count_out, int32_t *track_ptr_out, int32_t *track_det_out, void *stream) {
    M3_CHECK_ARG(n_graphs >= 0, "n_graphs < 0");
    if (n_graphs == 0) return MOT3D_OK;
    M3_CHECK_ARG(graph_ptr && payload && track_count_out && track_ptr_out && track_det_out, "NULL pointer");
    k_unpack_tracks<<<n_graphs, 256, 0, (cudaStream_t)stream>>>(n_graphs, graph_ptr, payload, track_count_out,
                                                               track_ptr_out, track_det_out);
    M3_LAUNCH_CHECK("k_unpack_tracks");
    return MOT3D_OK;
}

/* ---- peer buffers (CUDA IPC) ---------------------------------------------------------------------------------- */

int32_t mot3d_peer_buffer_create(size_t bytes, void **dev_ptr_out, unsigned char handle_out[MOT3D_PEER_HANDLE_BYTES]) {
    M3_CHECK_ARG(dev_ptr_out && handle_out && bytes > 0, "bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) <= MOT3D_PEER_HANDLE_BYTES, "handle size");
    void *p = nullptr;
    M3_CUDA(cudaMalloc(&p, bytes));
    cudaError_t e = cudaMemset(p, 0, bytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return cuda_fail(e, "cudaIpcGetMemHandle");
    }
    memset(handle_out, 0, MOT3D_PEER_HANDLE_BYTES);
    memcpy(handle_out, &h, sizeof(h));
    *dev_ptr_out = p;
    return MOT3D_OK;
}

int32_t mot3d_peer_buffer_open(const unsigned char handle[MOT3D_PEER_HANDLE_BYTES], void **dev_ptr_out) {
    M3_CHECK_ARG(handle && dev_ptr_out, "bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void *p = nullptr;
    M3_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *dev_ptr_out = p;
    return MOT3D_OK;
}

int32_t mot3d_peer_buffer_close(void *dev_ptr) {
    if (dev_ptr) M3_CUDA(cudaIpcCloseMemHandle(dev_ptr));
    return MOT3D_OK;
}

int32_t mot3d_peer_buffer_destroy(void *dev_ptr) {
    if (dev_ptr) M3_CUDA(cudaFree(dev_ptr));
    return MOT3D_OK;
}

int32_t mot3d_peer_put(void *peer_dst, const void *src, size_t bytes, int32_t *peer_flag, int32_t *local_flag_scratch,
                       int32_t step, void *stream) {
    M3_CHECK_ARG(peer_dst && src && peer_flag && local_flag_scratch, "NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    // stream order: the payload has landed before the flag is written
    if (bytes) M3_CUDA(cudaMemcpyAsync(peer_dst, src, bytes, cudaMemcpyDeviceToDevice, st));
    k_set_word<<<1, 1, 0, st>>>(local_flag_scratch, step);
    M3_LAUNCH_CHECK("k_set_word");
    M3_CUDA(cudaMemcpyAsync(peer_flag, local_flag_scratch, sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    return MOT3D_OK;
}

int32_t mot3d_peer_set_word(int32_t *word, int32_t value, void *stream) {
    M3_CHECK_ARG(word, "NULL pointer");
    k_set_word<<<1, 1, 0, (cudaStream_t)stream>>>(word, value);
    M3_LAUNCH_CHECK("k_set_word");
    return MOT3D_OK;
}

int32_t mot3d_peer_wait_word(const int32_t *word, int32_t need, int32_t *timeout_flag, void *stream) {
    M3_CHECK_ARG(word && timeout_flag, "NULL pointer");
    k_wait_word<<<1, 1, 0, (cudaStream_t)stream>>>(word, need, timeout_flag);
    M3_LAUNCH_CHECK("k_wait_word");
    return MOT3D_OK;
}

int32_t mot3d_peer_wait(const int32_t *flags, int32_t stride_words, int32_t world, int32_t step, int32_t *timeout_flag,
                        void *stream) {
    M3_CHECK_ARG(flags && timeout_flag && world >= 1 && stride_words >= 1, "bad arguments");
    if (world == 1) return MOT3D_OK;
    k_wait_flags<<<1, 1, 0, (cudaStream_t)stream>>>(flags, stride_words, world, step, timeout_flag);
    M3_LAUNCH_CHECK("k_wait_flags");
    return MOT3D_OK;
}

}  // extern "C"
