/* CPU restatement of the reference's native kNN + grouping path.
 *
 * TEST INFRASTRUCTURE ONLY -- never linked into, loaded by or called from the
 * product (gcanet_b200/).  Plain C, built by oracle/Makefile into
 * oracle/_build/liboracle_native.so.
 *
 * Follows (paths relative to /root/reference):
 *   models/KNN_CUDA/knn_cuda/csrc/cuda/knn.cu:29-93    cuComputeDistanceGlobal
 *       ssd = sum_d (ref[d][r] - query[d][q])^2, dims in ascending order; nvcc
 *       contracts `ssd += tmp*tmp` into one FMA, restated here with fmaf().
 *       The 16-wide zero padding of the dim loop adds exact zeros.
 *   models/KNN_CUDA/knn_cuda/csrc/cuda/knn.cu:105-167  cuInsertionSort
 *       per query: the k smallest, ascending; comparisons are strict (:125,:149)
 *       so among equal distances the lower reference index stays first;
 *       written indices are 1-based (:119,:138,:162).
 *   models/KNN_CUDA/knn_cuda/csrc/cuda/knn.cu:178-183  cuParallelSqrt
 *   models/Pointnet2_PyTorch-master/pointnet2_ops_lib/pointnet2_ops/_ext-src/src/
 *       group_points_gpu.cu:8-28 (gather), :43-64 (atomicAdd scatter).
 *
 * Parity pin: tests/test_oracle_golden.py checks this file against the golden
 * vectors of models/search_knn.py:180-304 and against sklearn's KDTree on the
 * size grid of models/KNN_CUDA/tests/test_knn_cuda.py:59-87 (distances, 3
 * decimals -- the reference's own criterion); on a GPU box the reference's
 * knn.cu itself (oracle/_ref/libknn_cuda_ref.so) is run beside it.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ref [dim][ref_nb], query [dim][query_nb] -> dist [k][query_nb], ind [k][query_nb]
 * (1-based, like knn_device).  Returns 0, or -1 on bad arguments. */
int oracle_knn_device(const float *ref, int ref_nb, const float *query, int query_nb,
                      int dim, int k, float *dist, int64_t *ind)
{
    if (k < 1 || k > ref_nb || dim < 1) return -1;
    float *best_d = (float *)malloc(sizeof(float) * (size_t)k);
    int64_t *best_i = (int64_t *)malloc(sizeof(int64_t) * (size_t)k);
    if (!best_d || !best_i) { free(best_d); free(best_i); return -1; }

    for (int q = 0; q < query_nb; ++q) {
        int filled = 0;
        for (int r = 0; r < ref_nb; ++r) {
            float ssd = 0.0f;
            for (int d = 0; d < dim; ++d) {
                float t = ref[(size_t)d * ref_nb + r] - query[(size_t)d * query_nb + q];
                ssd = fmaf(t, t, ssd);
            }
            if (filled == k && !(ssd < best_d[k - 1])) continue;
            /* slot after every entry that is <= ssd (stable for ties) */
            int pos = filled < k ? filled : k - 1;
            while (pos > 0 && best_d[pos - 1] > ssd) {
                best_d[pos] = best_d[pos - 1];
                best_i[pos] = best_i[pos - 1];
                --pos;
            }
            best_d[pos] = ssd;
            best_i[pos] = (int64_t)r + 1;
            if (filled < k) ++filled;
        }
        for (int j = 0; j < k; ++j) {
            dist[(size_t)j * query_nb + q] = sqrtf(best_d[j]);
            ind[(size_t)j * query_nb + q] = best_i[j];
        }
    }
    free(best_d);
    free(best_i);
    return 0;
}

/* points [b][c][n], idx [b][npoints][nsample] -> out [b][c][npoints][nsample] */
int oracle_group_points(int b, int c, int n, int npoints, int nsample,
                        const float *points, const int32_t *idx, float *out)
{
    for (int bi = 0; bi < b; ++bi)
        for (int l = 0; l < c; ++l)
            for (int j = 0; j < npoints; ++j)
                for (int s = 0; s < nsample; ++s) {
                    int32_t ii = idx[((size_t)bi * npoints + j) * nsample + s];
                    if (ii < 0 || ii >= n) return -1;
                    out[(((size_t)bi * c + l) * npoints + j) * nsample + s] =
                        points[((size_t)bi * c + l) * n + ii];
                }
    return 0;
}

/* grad_out [b][c][npoints][nsample] -> grad_points [b][c][n] (zeroed here) */
int oracle_group_points_grad(int b, int c, int n, int npoints, int nsample,
                             const float *grad_out, const int32_t *idx, float *grad_points)
{
    memset(grad_points, 0, sizeof(float) * (size_t)b * c * n);
    for (int bi = 0; bi < b; ++bi)
        for (int l = 0; l < c; ++l)
            for (int j = 0; j < npoints; ++j)
                for (int s = 0; s < nsample; ++s) {
                    int32_t ii = idx[((size_t)bi * npoints + j) * nsample + s];
                    if (ii < 0 || ii >= n) return -1;
                    grad_points[((size_t)bi * c + l) * n + ii] +=
                        grad_out[(((size_t)bi * c + l) * npoints + j) * nsample + s];
                }
    return 0;
}

/* ---------------------------------------------------------------------------------------------------------
 * Gated ball query: restates ballquery_batch_p_cuda_ (softgroup/ops/src/bfs_cluster/bfs_cluster.cu:18-77).
 * One "thread" per point, in point order: candidates k of the point's own batch segment in ascending order,
 * kept iff d2 < radius^2 (same expression order, no contraction: built with -ffp-contract=off) and both dense
 * affinities exceed their thresholds; at most 3000 per point (the kernel's idx_temp[3000] stops the scan at the
 * 3001st hit); lists are concatenated in point order (the kernel: in the order its atomicAdd happens to run).
 * start_len [n][2] = (start, length); returns the total number of neighbours, idx must hold that many (call once
 * with idx == NULL to size it).
 * Pinned: oracle/make_golden.py runs the reference kernel's own text on the host (same point order) and requires
 * identical outputs, including the 3000 cap and the retry loop of functions.py:460-475; fixture
 * tests/golden/ballquery_small.npz. */
long long oracle_ballquery_batch_p(int n, float radius, const float *xyz, const int32_t *batch_idxs,
                                   const int32_t *batch_offsets, const float *adj_inst, float thr_inst,
                                   const float *adj_para, float thr_para, int32_t *idx, int32_t *start_len)
{
    const float radius2 = radius * radius;
    long long total = 0;
    for (int p = 0; p < n; ++p) {
        const float ox = xyz[p * 3 + 0], oy = xyz[p * 3 + 1], oz = xyz[p * 3 + 2];
        const int b = batch_idxs[p];
        const int start = batch_offsets[b], end = batch_offsets[b + 1];
        int cnt = 0;
        for (int k = start; k < end; ++k) {
            const float x = xyz[k * 3 + 0], y = xyz[k * 3 + 1], z = xyz[k * 3 + 2];
            const float d2 = (ox - x) * (ox - x) + (oy - y) * (oy - y) + (oz - z) * (oz - z);
            if (d2 < radius2 && adj_inst[(size_t)p * n + k] > thr_inst && adj_para[(size_t)p * n + k] > thr_para) {
                if (cnt < 3000) {
                    if (idx) idx[total + cnt] = k;
                } else {
                    break;
                }
                ++cnt;
            }
        }
        if (start_len) { start_len[p * 2] = (int32_t)total; start_len[p * 2 + 1] = cnt; }
        total += cnt;
    }
    return total;
}
