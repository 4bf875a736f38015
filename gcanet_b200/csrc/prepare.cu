// Input side of the hot path: what ABCDataset.__getitem__ does to a raw sample after reading it
// (dataloader/ABCDataset_new.py:77-141), getInstanceInfo (:157-178) and the stacking of collate_fn (:182-295), for a whole
// batch on the device -- so that raw shards can be kept in HBM and a step's batch is produced next to the kernels that
// consume it, already in both layouts they want: channel-major [B][6][n] (the reference's, train_new.py:25-26) and
// point-major [B][n][8] (xyz | normal | 0 0, what the fused EdgeConv kernels read).
//
// One CTA per cloud:
//   1. histogram of the raw instance labels + position of each label's first appearance (shared-memory atomics);
//   2. instances with more than `min_points` raw points are kept and renumbered in order of first appearance (the
//      reference iterates a Counter, i.e. insertion order), the rest become background (-1);
//   3. the subsample (indices drawn by the caller, like np.random.choice on the host) is gathered; T_gt = primitive type of
//      kept instances else -1 (7-class remap 7 -> 6, 9 -> 6, 8 -> 2), I_gt_clean = new id or old label + number kept;
//   4. per kept instance: mean of its subsampled points (fp64 accumulation), size, type of its first point;
//      pt_offset_label = mean - point, with -100 standing in for the mean of background points.
// Integer outputs are bit-exact against the reference; the offsets agree to fp32 rounding of the mean.
#include "common.cuh"

#include <limits.h>

namespace gcanet {

struct PrepArgs {
    const float *points, *normals, *t_param;
    const int32_t *labels, *prim, *sub;
    float *cloud_cn, *cloud_nc, *t_param_out, *pt_offset;
    int32_t *i_gt, *t_gt, *i_clean, *inst_num, *inst_pointnum, *inst_cls, *status;
    int n_raw, n_sub, L, min_points, num_primitives;
};

__global__ void __launch_bounds__(1024) prepare_samples_kernel(PrepArgs a) {
    extern __shared__ __align__(8) unsigned char prep_sm[];
    double *s_sum = reinterpret_cast<double *>(prep_sm);            // [L][3]
    int *s_count = reinterpret_cast<int *>(s_sum + 3 * a.L);        // [L] raw points per label
    int *s_first = s_count + a.L;                                   // [L] first raw position
    int *s_map = s_first + a.L;                                     // [L] new id or -1
    int *s_icnt = s_map + a.L;                                      // [L] subsampled points per new id
    int *s_ifirst = s_icnt + a.L;                                   // [L] first subsample position per new id
    __shared__ int s_nkeys, s_bad, s_imax;
    const int b = blockIdx.x, t = threadIdx.x, nt = blockDim.x;
    const int32_t *lab = a.labels + (size_t)b * a.n_raw;
    for (int l = t; l < a.L; l += nt) {
        s_count[l] = 0; s_first[l] = INT_MAX; s_icnt[l] = 0; s_ifirst[l] = INT_MAX;
        s_sum[3 * l] = 0.0; s_sum[3 * l + 1] = 0.0; s_sum[3 * l + 2] = 0.0;
    }
    if (t == 0) { s_nkeys = 0; s_bad = 0; s_imax = -1; }
    __syncthreads();
    for (int n = t; n < a.n_raw; n += nt) {
        const int l = lab[n];
        if (l < 0 || l >= a.L) { s_bad = 1; continue; }
        atomicAdd(&s_count[l], 1);
        atomicMin(&s_first[l], n);
    }
    __syncthreads();
    if (s_bad) {                                                    // label outside [0, max_labels): report, write nothing
        if (t == 0) { a.status[b] = 1; a.inst_num[b] = 0; }
        return;
    }
    for (int l = t; l < a.L; l += nt) {
        int m = -1;
        if (s_count[l] > a.min_points) {
            m = 0;
            for (int o = 0; o < a.L; ++o) m += (s_count[o] > a.min_points && s_first[o] < s_first[l]) ? 1 : 0;
            atomicAdd(&s_nkeys, 1);
        }
        s_map[l] = m;
    }
    __syncthreads();
    const int nkeys = s_nkeys;
    const size_t sb = (size_t)b * a.n_sub;
    for (int s = t; s < a.n_sub; s += nt) {
        int r = a.sub[sb + s];
        r = r < 0 ? 0 : (r >= a.n_raw ? a.n_raw - 1 : r);
        const size_t rr = (size_t)b * a.n_raw + r;
        const int l = lab[r], m = s_map[l];
        int tg = m >= 0 ? a.prim[rr] : -1;
        if (a.num_primitives == 7) tg = (tg == 7 || tg == 9) ? 6 : (tg == 8 ? 2 : tg);
        a.i_gt[sb + s] = m;
        a.t_gt[sb + s] = tg;
        a.i_clean[sb + s] = m >= 0 ? m : l + nkeys;
        const float px = a.points[rr * 3], py = a.points[rr * 3 + 1], pz = a.points[rr * 3 + 2];
        const float nx = a.normals[rr * 3], ny = a.normals[rr * 3 + 1], nz = a.normals[rr * 3 + 2];
        float *cn = a.cloud_cn + (size_t)b * 6 * a.n_sub + s;
        cn[0] = px; cn[(size_t)a.n_sub] = py; cn[2 * (size_t)a.n_sub] = pz;
        cn[3 * (size_t)a.n_sub] = nx; cn[4 * (size_t)a.n_sub] = ny; cn[5 * (size_t)a.n_sub] = nz;
        float4 *nc = reinterpret_cast<float4 *>(a.cloud_nc + (sb + s) * 8);
        nc[0] = make_float4(px, py, pz, nx);
        nc[1] = make_float4(ny, nz, 0.f, 0.f);
        for (int c = 0; c < 22; ++c) a.t_param_out[(sb + s) * 22 + c] = a.t_param[rr * 22 + c];
        if (m >= 0) {
            atomicAdd(&s_sum[3 * m], (double)px);
            atomicAdd(&s_sum[3 * m + 1], (double)py);
            atomicAdd(&s_sum[3 * m + 2], (double)pz);
            atomicAdd(&s_icnt[m], 1);
            atomicMin(&s_ifirst[m], s);
            atomicMax(&s_imax, m);
        }
    }
    __syncthreads();
    const int inum = s_imax + 1;                                    // max(instance_label) + 1 over the subsample (:162)
    for (int s = t; s < a.n_sub; s += nt) {
        const int m = a.i_gt[sb + s];
        float mx = -100.f, my = -100.f, mz = -100.f;
        if (m >= 0) {
            const double c = (double)s_icnt[m];
            mx = (float)(s_sum[3 * m] / c); my = (float)(s_sum[3 * m + 1] / c); mz = (float)(s_sum[3 * m + 2] / c);
        }
        const float *p = a.cloud_nc + (sb + s) * 8;
        a.pt_offset[(sb + s) * 3] = mx - p[0];
        a.pt_offset[(sb + s) * 3 + 1] = my - p[1];
        a.pt_offset[(sb + s) * 3 + 2] = mz - p[2];
    }
    for (int i = t; i < a.L; i += nt) {
        const bool live = i < inum && s_icnt[i] > 0;
        a.inst_pointnum[(size_t)b * a.L + i] = i < inum ? s_icnt[i] : 0;
        a.inst_cls[(size_t)b * a.L + i] = live ? a.t_gt[sb + s_ifirst[i]] : -1;
    }
    if (t == 0) { a.inst_num[b] = inum; a.status[b] = 0; }
}

}  // namespace gcanet

using namespace gcanet;

extern "C" int gcanet_prepare_samples(const gcanet_prepare_desc *d, const float *points, const float *normals, const int32_t *labels,
                                      const int32_t *prim, const float *t_param, const int32_t *sub_index, float *cloud_cn,
                                      float *cloud_nc, int32_t *i_gt, int32_t *t_gt, int32_t *i_gt_clean, float *t_param_out,
                                      float *pt_offset_label, int32_t *inst_num, int32_t *inst_pointnum, int32_t *inst_cls,
                                      int32_t *status, gcanet_stream_t stream) {
    GCANET_REQUIRE(d != nullptr, "prepare_samples: null descriptor");
    GCANET_REQUIRE(d->B >= 1 && d->B <= 65535 && d->n_raw >= 1 && d->n_sub >= 1, "prepare_samples: bad shape B=%d n_raw=%d n_sub=%d",
                   d->B, d->n_raw, d->n_sub);
    GCANET_REQUIRE(d->max_labels >= 1 && d->max_labels <= 4096, "prepare_samples: max_labels=%d must be in [1, 4096]", d->max_labels);
    GCANET_REQUIRE(d->min_points >= 0, "prepare_samples: min_points < 0");
    GCANET_REQUIRE(points && normals && labels && prim && t_param && sub_index && cloud_cn && cloud_nc && i_gt && t_gt && i_gt_clean &&
                   t_param_out && pt_offset_label && inst_num && inst_pointnum && inst_cls && status, "prepare_samples: null pointer");
    GCANET_REQUIRE((reinterpret_cast<uintptr_t>(cloud_nc) & 15) == 0, "prepare_samples: cloud_nc must be 16-byte aligned");
    PrepArgs a{points, normals, t_param, labels, prim, sub_index, cloud_cn, cloud_nc, t_param_out, pt_offset_label,
               i_gt, t_gt, i_gt_clean, inst_num, inst_pointnum, inst_cls, status, d->n_raw, d->n_sub, d->max_labels, d->min_points,
               d->num_primitives};
    const size_t smem = (size_t)d->max_labels * (3 * sizeof(double) + 5 * sizeof(int));
    GCANET_CUDA_OK(cudaFuncSetAttribute(prepare_samples_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    prepare_samples_kernel<<<d->B, 1024, smem, as_stream(stream)>>>(a);
    GCANET_LAUNCH_OK("prepare_samples_kernel");
    return GCANET_OK;
}
