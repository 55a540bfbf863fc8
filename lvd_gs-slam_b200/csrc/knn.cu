// knn.cu -- K12: simple_knn.distCUDA2 replacement (reached from
// GaussianModel.extend_from_pcd_seq via utils/slam_backend.py:75-78): for every point the mean of the squared
// distances to its three nearest neighbours, exact.
//
// Round-1 implementation: all-pairs scan tiled through shared memory (one query per thread, 256 candidates per
// stage as float4).  Exact by construction; O(P^2) FP32 work, which at the reference's per-keyframe sizes
// (7.3k-14.6k points, configs/mono/KITTI/base_config.yaml:16-17) is a few microseconds of B200 time.
#include "common.cuh"
#include <float.h>

namespace lvdgs {

constexpr int KNN_THREADS = 256;

__global__ void __launch_bounds__(KNN_THREADS) dist2_bruteforce_kernel(int P, const float *__restrict__ pts,
                                                                       float *__restrict__ out) {
    __shared__ float4 s_p[KNN_THREADS];
    const int i = blockIdx.x * KNN_THREADS + threadIdx.x;
    float3 q = make_float3(0.f, 0.f, 0.f);
    if (i < P) q = make_float3(pts[3 * (size_t)i], pts[3 * (size_t)i + 1], pts[3 * (size_t)i + 2]);
    float b0 = FLT_MAX, b1 = FLT_MAX, b2 = FLT_MAX;
    for (int base = 0; base < P; base += KNN_THREADS) {
        const int j = base + threadIdx.x;
        __syncthreads();
        s_p[threadIdx.x] = j < P ? make_float4(pts[3 * (size_t)j], pts[3 * (size_t)j + 1], pts[3 * (size_t)j + 2], 0.f)
                                 : make_float4(FLT_MAX, FLT_MAX, FLT_MAX, 0.f);
        __syncthreads();
        const int nb = min(KNN_THREADS, P - base);
#pragma unroll 8
        for (int k = 0; k < nb; ++k) {
            const float4 c = s_p[k];
            const float dx = c.x - q.x, dy = c.y - q.y, dz = c.z - q.z;
            const float d = dx * dx + dy * dy + dz * dz;
            if (d < b2 && base + k != i) {
                if (d < b1) {
                    b2 = b1;
                    if (d < b0) { b1 = b0; b0 = d; } else b1 = d;
                } else b2 = d;
            }
        }
    }
    if (i < P) out[i] = (b0 + b1 + b2) / 3.f;
}

size_t dist2_workspace_bytes(int P) { (void)P; return 256; }

int launch_dist2(int P, const float *points, float *mean_dists, void *ws, size_t ws_bytes, cudaStream_t s) {
    (void)ws; (void)ws_bytes;
    if (P <= 0) return 0;
    LVDGS_PRE(s);
    dist2_bruteforce_kernel<<<ceil_div(P, KNN_THREADS), KNN_THREADS, 0, s>>>(P, points, mean_dists);
    LVDGS_LAUNCHED(s, "dist2");
    return 0;
}

}  // namespace lvdgs
