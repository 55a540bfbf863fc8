// blend_forward.cu -- K6: per-16x16-tile front-to-back alpha blending of colour(3) + depth, with the
// opacity (1 - T) and n_touched outputs of the pose-aware fork (reached from
// utils/slam_frontend.py:1493 / utils/slam_backend.py:184 through gaussian_renderer.render).
//
// One CTA per tile, one thread per pixel.  The tile's depth-sorted instance list is streamed in batches of 256:
// each thread gathers one instance's 40 bytes (xy, conic+opacity, rgb+depth -- three aligned vector loads from the
// SoA geometry arrays, which are L2-resident) into shared memory, then all threads walk the batch with broadcast
// shared-memory reads.  FP32 FMA/MUFU bound; no tensor cores (there is no dense contraction on this path).
// n_touched is aggregated per warp with a ballot, and only while some pixel of the warp still has T > 0.5
// (T only decreases, so the test T*(1-alpha) > 0.5 can never fire afterwards).
#include "common.cuh"

namespace lvdgs {

constexpr int BF_THREADS = TILE_PIX;

__global__ void __launch_bounds__(BF_THREADS) blend_forward_kernel(int W, int H, int gx, const uint2 *__restrict__ ranges,
                                                                   const uint32_t *__restrict__ point_list,
                                                                   const float2 *__restrict__ means2D,
                                                                   const float4 *__restrict__ conic_opacity,
                                                                   const float4 *__restrict__ rgbd, const float *__restrict__ bg,
                                                                   float *__restrict__ out_color, float *__restrict__ out_depth,
                                                                   float *__restrict__ out_opacity, float *__restrict__ final_T,
                                                                   uint32_t *__restrict__ n_contrib, int32_t *__restrict__ n_touched) {
    __shared__ uint32_t s_id[BF_THREADS];
    __shared__ float2 s_xy[BF_THREADS];
    __shared__ float4 s_co[BF_THREADS];
    __shared__ float4 s_cd[BF_THREADS];

    const int tile = blockIdx.y * gx + blockIdx.x;
    const int lx = threadIdx.x & (TILE - 1), ly = threadIdx.x >> 4;
    const int px = blockIdx.x * TILE + lx, py = blockIdx.y * TILE + ly;
    const bool inside = px < W && py < H;
    const float pfx = (float)px, pfy = (float)py;
    const int lane = threadIdx.x & 31;

    const uint2 range = ranges[tile];
    int todo = (int)(range.y - range.x);
    bool done = !inside;
    float T = 1.f, C0 = 0.f, C1 = 0.f, C2 = 0.f, D = 0.f;
    uint32_t contributor = 0, last_contributor = 0;
    bool warp_hi = true;     // some pixel of this warp may still satisfy T(1-alpha) > 0.5

    for (uint32_t base = range.x; todo > 0; base += BF_THREADS, todo -= BF_THREADS) {
        if (__syncthreads_count(done) == BF_THREADS) break;
        if ((int)threadIdx.x < todo) {
            const uint32_t id = __ldg(point_list + base + threadIdx.x);
            s_id[threadIdx.x] = id;
            s_xy[threadIdx.x] = __ldg(means2D + id);
            s_co[threadIdx.x] = __ldg(conic_opacity + id);
            s_cd[threadIdx.x] = __ldg(rgbd + id);
        }
        __syncthreads();
        const int nb = min(BF_THREADS, todo);
        for (int j = 0; j < nb; ++j) {
            if ((j & 7) == 0) {
                if (__all_sync(0xffffffffu, done)) break;
                if (warp_hi) warp_hi = __any_sync(0xffffffffu, !done && T > 0.5f);
            }
            bool hit = false;
            if (!done) {
                contributor++;
                const float2 xy = s_xy[j];
                const float4 co = s_co[j];
                const float dx = xy.x - pfx, dy = xy.y - pfy;
                const float power = -0.5f * (co.x * dx * dx + co.z * dy * dy) - co.y * dx * dy;
                if (power <= 0.f) {
                    const float alpha = fminf(0.99f, co.w * __expf(power));
                    if (alpha >= 1.f / 255.f) {
                        const float test_T = T * (1.f - alpha);
                        if (test_T < 0.0001f) {
                            done = true;
                        } else {
                            const float4 cd = s_cd[j];
                            const float wgt = alpha * T;
                            C0 += cd.x * wgt; C1 += cd.y * wgt; C2 += cd.z * wgt; D += cd.w * wgt;
                            hit = test_T > 0.5f;
                            T = test_T;
                            last_contributor = contributor;
                        }
                    }
                }
            }
            if (warp_hi) {
                const uint32_t b = __ballot_sync(0xffffffffu, hit);
                if (b && lane == 0) atomicAdd(n_touched + s_id[j], __popc(b));
            }
        }
    }
    if (inside) {
        const size_t pix = (size_t)py * W + px;
        const size_t HW = (size_t)H * W;
        final_T[pix] = T;
        n_contrib[pix] = last_contributor;
        out_color[pix] = C0 + T * __ldg(bg + 0);
        out_color[HW + pix] = C1 + T * __ldg(bg + 1);
        out_color[2 * HW + pix] = C2 + T * __ldg(bg + 2);
        out_depth[pix] = D;
        out_opacity[pix] = 1.f - T;
    }
}

int launch_blend_forward(int W, int H, const uint2 *ranges, const uint32_t *point_list, const GeomPtrs &g,
                         const float *bg, float *out_color, float *out_depth, float *out_opacity, float *final_T,
                         uint32_t *n_contrib, int32_t *n_touched, cudaStream_t s) {
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    LVDGS_PRE(s);
    blend_forward_kernel<<<dim3(gx, gy), BF_THREADS, 0, s>>>(W, H, gx, ranges, point_list, g.means2D, g.conic_opacity,
                                                              g.rgbd, bg, out_color, out_depth, out_opacity, final_T,
                                                              n_contrib, n_touched);
    LVDGS_LAUNCHED(s, "blend_forward");
    return 0;
}

}  // namespace lvdgs
