// Input-side feature prep: merge_close_points, the per-frame de-duplication of detections the
// reference's 3-D example runs before build_graph (mot3d/utils/filt.py:6-24, examples/soldiers_3d/example.py:29-31).
//
// Reference: DBSCAN(eps, min_samples=1, metric='euclidean', algorithm='brute') = the connected components of the
// graph "distance <= eps"; clusters are numbered in the order of their first member; every cluster is replaced by the
// mean of its members (np.mean over the members in index order). One block per point set (frame): union-find in
// shared memory with roots hooked under the smaller root, so a cluster's root is its first member.
// (sklearn evaluates the squared distance as x.x - 2 x.y + y.y; the direct sum of squares used here can only differ
// for pairs whose distance equals eps to 1 part in 1e12.)
#include "common.cuh"

namespace m3 {

constexpr int MC_TPB = 128;
constexpr int MC_MAX_POINTS = 8192;  // per set (two int arrays in shared memory)

__global__ void __launch_bounds__(MC_TPB) k_merge_close_points(const int32_t *__restrict__ set_ptr, int dim,
                                                               const double *__restrict__ pts, double eps_sq,
                                                               int32_t *__restrict__ label, int32_t *__restrict__ count,
                                                               double *__restrict__ merged) {
    extern __shared__ int mc_smem[];
    __shared__ int s_changed;
    const int s = blockIdx.x;
    const int a = set_ptr[s], m = set_ptr[s + 1] - a;
    int *parent = mc_smem, *rank = mc_smem + m;
    const int tid = threadIdx.x;
    for (int i = tid; i < m; i += MC_TPB) parent[i] = i;
    __syncthreads();
    auto find = [&](int i) {
        while (true) {
            const int pi = parent[i];
            if (pi == i) return i;
            i = pi;
        }
    };
    const double *P = pts + (size_t)a * dim;
    const long long pairs = (long long)m * (m - 1) / 2;
    for (int sweep = 0; sweep < m; sweep++) {
        if (tid == 0) s_changed = 0;
        __syncthreads();
        for (long long k = tid; k < pairs; k += MC_TPB) {
            // pair (i, j), i < j, from the linear index k = j (j - 1) / 2 + i
            int j = (int)((1.0 + sqrt(1.0 + 8.0 * (double)k)) * 0.5);
            while ((long long)j * (j - 1) / 2 > k) j--;
            while ((long long)(j + 1) * j / 2 <= k) j++;
            const int i = (int)(k - (long long)j * (j - 1) / 2);
            double d2 = 0.0;
            for (int c = 0; c < dim; c++) {
                const double e = P[(size_t)i * dim + c] - P[(size_t)j * dim + c];
                d2 += e * e;
            }
            if (d2 <= eps_sq) {
                const int ri = find(i), rj = find(j);
                if (ri != rj) {
                    atomicMin(&parent[ri > rj ? ri : rj], ri < rj ? ri : rj);  // hook the larger root under the smaller
                    s_changed = 1;
                }
            }
        }
        __syncthreads();
        if (!s_changed) break;
        __syncthreads();
    }
    // roots, cluster numbers by first member, means over the members in index order
    for (int i = tid; i < m; i += MC_TPB) rank[i] = find(i);
    __syncthreads();
    for (int i = tid; i < m; i += MC_TPB) parent[i] = rank[i];  // fully compressed
    __syncthreads();
    for (int i = tid; i < m; i += MC_TPB) {
        const int r = parent[i];
        int c = 0;
        for (int q = 0; q < r; q++) c += parent[q] == q ? 1 : 0;
        label[a + i] = c;
        if (r == i) {
            double acc[3] = {0.0, 0.0, 0.0};
            int k = 0;
            for (int q = i; q < m; q++)
                if (parent[q] == i) {
                    for (int d = 0; d < dim; d++) acc[d] = acc[d] + P[(size_t)q * dim + d];
                    k++;
                }
            for (int d = 0; d < dim; d++) merged[((size_t)a + c) * dim + d] = acc[d] / (double)k;
        }
    }
    if (tid == 0) {
        int c = 0;
        for (int q = 0; q < m; q++) c += parent[q] == q ? 1 : 0;
        count[s] = c;
    }
}

}  // namespace m3

using namespace m3;

extern "C" {

int32_t mot3d_merge_close_points(const int32_t *set_ptr, int32_t n_sets, int32_t max_set_points, int32_t dim,
                                 const double *points, double eps, int32_t *label_out, int32_t *cluster_count_out,
                                 double *merged_out, void *stream) {
    M3_CHECK_ARG(n_sets >= 0 && max_set_points >= 0, "negative sizes");
    if (n_sets == 0) return MOT3D_OK;
    M3_CHECK_ARG(dim >= 1 && dim <= 3, "dim must be 1, 2 or 3");
    M3_CHECK_ARG(set_ptr && label_out && cluster_count_out && merged_out, "NULL pointer");
    M3_CHECK_ARG(max_set_points == 0 || points, "points is NULL");
    if (max_set_points > MC_MAX_POINTS) {
        set_error("mot3d_merge_close_points: %d points in one set, at most %d are supported", max_set_points,
                  MC_MAX_POINTS);
        return MOT3D_E_UNSUPPORTED;
    }
    const size_t smem = (size_t)(max_set_points > 0 ? max_set_points : 1) * 2 * sizeof(int);
    if (smem > 48 * 1024)
        M3_CUDA(cudaFuncSetAttribute(k_merge_close_points, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_merge_close_points<<<n_sets, MC_TPB, smem, (cudaStream_t)stream>>>(set_ptr, dim, points, eps * eps, label_out,
                                                                         cluster_count_out, merged_out);
    M3_LAUNCH_CHECK("k_merge_close_points");
    return MOT3D_OK;
}

}  // extern "C"
