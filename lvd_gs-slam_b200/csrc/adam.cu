// adam.cu -- fused Adam update of the replicated parameter block after the gradient all-reduce of the keyframe-
// sharded mapping iteration (lvdgs.mapping; the reference calls torch.optim.Adam.step on five parameter groups,
// utils/slam_backend.py:378-380 via GaussianModel.optimizer).  One pass over the [P,14] block: 16 B read (p, g, m, v)
// + 12 B written (p, m, v) per element, HBM-bound; per-group learning rates are resolved from the group offsets.
#include "common.cuh"
#include <cmath>

namespace lvdgs {

struct AdamArgs {
    int64_t n;
    int64_t group_end[8];     // exclusive end offset of each group
    float lr[8];
    int groups;
    float b1, b2, omb1, omb2, eps, inv_bc1, inv_sqrt_bc2;
};

__global__ void __launch_bounds__(256) adam_step_kernel(float *__restrict__ p, const float *__restrict__ g,
                                                        float *__restrict__ m, float *__restrict__ v, const AdamArgs a) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride) {
        float lr = a.lr[0];
#pragma unroll
        for (int k = 1; k < 8; ++k)
            if (k < a.groups && i >= a.group_end[k - 1]) lr = a.lr[k];
        const float gi = g[i];
        const float mi = a.b1 * m[i] + a.omb1 * gi;
        const float vi = a.b2 * v[i] + a.omb2 * gi * gi;
        m[i] = mi; v[i] = vi;
        const float denom = sqrtf(vi) * a.inv_sqrt_bc2 + a.eps;
        p[i] -= lr * a.inv_bc1 * mi / denom;
    }
}

int launch_adam_step(int64_t n, float *params, const float *grads, float *exp_avg, float *exp_avg_sq, int groups,
                     const int64_t *group_end, const float *lr, double beta1, double beta2, double eps, int step,
                     cudaStream_t s) {
    if (n <= 0) return 0;
    if (groups < 1 || groups > 8) { set_error("adam: 1..8 groups"); return 1; }
    AdamArgs a;
    a.n = n; a.groups = groups;
    for (int k = 0; k < 8; ++k) { a.group_end[k] = k < groups ? group_end[k] : n; a.lr[k] = k < groups ? lr[k] : 0.f; }
    a.b1 = (float)beta1; a.b2 = (float)beta2; a.eps = (float)eps;
    a.omb1 = (float)(1.0 - beta1); a.omb2 = (float)(1.0 - beta2);
    a.inv_bc1 = (float)(1.0 / (1.0 - pow(beta1, (double)step)));          // bias corrections in double, like torch
    a.inv_sqrt_bc2 = (float)(1.0 / sqrt(1.0 - pow(beta2, (double)step)));
    const int blocks = (int)min((int64_t)148 * 16, (n + 255) / 256);
    LVDGS_PRE(s);
    adam_step_kernel<<<blocks, 256, 0, s>>>(params, grads, exp_avg, exp_avg_sq, a);
    LVDGS_LAUNCHED(s, "adam_step");
    return 0;
}

}  // namespace lvdgs
