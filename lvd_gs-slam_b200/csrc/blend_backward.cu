// blend_backward.cu -- K7: back-to-front gradient of the tile blend (reached through
// loss.backward() at utils/slam_frontend.py:1517 and utils/slam_backend.py:306).
//
// The reference issues ten global float atomicAdds per (pixel, Gaussian) pair.  Here a CTA owns one tile and a
// thread one pixel, the per-pixel recurrences (T, accumulated colour/depth behind the current Gaussian) are
// those of A.4, but the ten per-Gaussian partial sums are first reduced across the warp with shuffles
// (skipped entirely when no pixel of the warp was touched by the Gaussian) and committed by one lane into a
// 48-byte accumulator row per Gaussian.  The traversal starts at the tile's largest n_contrib, not at the end
// of the tile's list, so the part of the list that every pixel terminated before is never loaded.
#include "common.cuh"

namespace lvdgs {

constexpr int BB_THREADS = TILE_PIX;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

__global__ void __launch_bounds__(BB_THREADS) blend_backward_kernel(
    int W, int H, int gx, const uint2 *__restrict__ ranges, const uint32_t *__restrict__ point_list,
    const float2 *__restrict__ means2D, const float4 *__restrict__ conic_opacity, const float4 *__restrict__ rgbd,
    const float *__restrict__ bg, const float *__restrict__ final_T, const uint32_t *__restrict__ n_contrib,
    const float *__restrict__ dL_dout_color, const float *__restrict__ dL_dout_depth,
    const float *__restrict__ dL_dout_opacity, float *__restrict__ acc) {
    __shared__ uint32_t s_id[BB_THREADS];
    __shared__ float2 s_xy[BB_THREADS];
    __shared__ float4 s_co[BB_THREADS];
    __shared__ float4 s_cd[BB_THREADS];
    __shared__ uint32_t s_max[BB_THREADS / 32];

    const int tile = blockIdx.y * gx + blockIdx.x;
    const int lx = threadIdx.x & (TILE - 1), ly = threadIdx.x >> 4;
    const int px = blockIdx.x * TILE + lx, py = blockIdx.y * TILE + ly;
    const bool inside = px < W && py < H;
    const float pfx = (float)px, pfy = (float)py;
    const int lane = threadIdx.x & 31;
    const size_t pix = (size_t)py * W + px, HW = (size_t)H * W;

    const uint2 range = ranges[tile];
    const uint32_t last = inside ? n_contrib[pix] : 0u;   // this pixel handles contributor indices < last
    const float T_final = inside ? final_T[pix] : 0.f;
    float T = T_final;
    float dpx0 = 0.f, dpx1 = 0.f, dpx2 = 0.f, dpd = 0.f, bg_dot = 0.f;
    if (inside) {
        dpx0 = dL_dout_color[pix]; dpx1 = dL_dout_color[HW + pix]; dpx2 = dL_dout_color[2 * HW + pix];
        if (dL_dout_depth) dpd = dL_dout_depth[pix];
        bg_dot = __ldg(bg) * dpx0 + __ldg(bg + 1) * dpx1 + __ldg(bg + 2) * dpx2;
        if (dL_dout_opacity) bg_dot -= dL_dout_opacity[pix];
    }
    // tile-wide max of n_contrib: nothing beyond it contributes to any pixel
    uint32_t m = last;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, d));
    if (lane == 0) s_max[threadIdx.x >> 5] = m;
    __syncthreads();
    uint32_t top = 0;
#pragma unroll
    for (int k = 0; k < BB_THREADS / 32; ++k) top = max(top, s_max[k]);
    top = min(top, range.y - range.x);

    float a0 = 0.f, a1 = 0.f, a2 = 0.f, ad = 0.f;           // accum_rec colour / depth
    float last_alpha = 0.f, lc0 = 0.f, lc1 = 0.f, lc2 = 0.f, ld = 0.f;
    const float half_W = 0.5f * (float)W, half_H = 0.5f * (float)H;

    // entries are visited in decreasing contributor index k = top-1 ... 0
    for (int remaining = (int)top; remaining > 0; remaining -= BB_THREADS) {
        __syncthreads();
        const int nb = min(BB_THREADS, remaining);
        if ((int)threadIdx.x < nb) {
            const uint32_t id = __ldg(point_list + range.x + (uint32_t)(remaining - 1 - (int)threadIdx.x));
            s_id[threadIdx.x] = id;
            s_xy[threadIdx.x] = __ldg(means2D + id);
            s_co[threadIdx.x] = __ldg(conic_opacity + id);
            s_cd[threadIdx.x] = __ldg(rgbd + id);
        }
        __syncthreads();
        for (int j = 0; j < nb; ++j) {
            const uint32_t k = (uint32_t)(remaining - 1 - j);    // contributor index (0-based) of this entry
            float g_mx = 0.f, g_my = 0.f, g_cxx = 0.f, g_cxy = 0.f, g_cyy = 0.f, g_op = 0.f, g_dd = 0.f,
                  g_c0 = 0.f, g_c1 = 0.f, g_c2 = 0.f;
            bool valid = false;
            if (k < last) {
                const float2 xy = s_xy[j];
                const float4 co = s_co[j];
                const float dx = xy.x - pfx, dy = xy.y - pfy;
                const float power = -0.5f * (co.x * dx * dx + co.z * dy * dy) - co.y * dx * dy;
                if (power <= 0.f) {
                    const float G = __expf(power);
                    const float alpha = fminf(0.99f, co.w * G);
                    if (alpha >= 1.f / 255.f) {
                        valid = true;
                        const float4 cd = s_cd[j];
                        const float one_m = 1.f - alpha;
                        T = __fdividef(T, one_m);
                        const float wgt = alpha * T;
                        float dL_dalpha;
                        a0 = last_alpha * lc0 + (1.f - last_alpha) * a0; lc0 = cd.x;
                        a1 = last_alpha * lc1 + (1.f - last_alpha) * a1; lc1 = cd.y;
                        a2 = last_alpha * lc2 + (1.f - last_alpha) * a2; lc2 = cd.z;
                        ad = last_alpha * ld + (1.f - last_alpha) * ad; ld = cd.w;
                        dL_dalpha = (cd.x - a0) * dpx0 + (cd.y - a1) * dpx1 + (cd.z - a2) * dpx2 + (cd.w - ad) * dpd;
                        g_c0 = wgt * dpx0; g_c1 = wgt * dpx1; g_c2 = wgt * dpx2; g_dd = wgt * dpd;
                        dL_dalpha *= T;
                        last_alpha = alpha;
                        dL_dalpha += __fdividef(-T_final, one_m) * bg_dot;
                        const float dL_dG = co.w * dL_dalpha;
                        const float gdx = G * dx, gdy = G * dy;
                        const float dG_ddelx = -gdx * co.x - gdy * co.y;
                        const float dG_ddely = -gdy * co.z - gdx * co.y;
                        g_mx = dL_dG * dG_ddelx * half_W;
                        g_my = dL_dG * dG_ddely * half_H;
                        g_cxx = -0.5f * gdx * dx * dL_dG;
                        g_cxy = -0.5f * gdx * dy * dL_dG;
                        g_cyy = -0.5f * gdy * dy * dL_dG;
                        g_op = G * dL_dalpha;
                    }
                }
            }
            if (__any_sync(0xffffffffu, valid)) {
                g_mx = warp_sum(g_mx); g_my = warp_sum(g_my);
                g_cxx = warp_sum(g_cxx); g_cxy = warp_sum(g_cxy); g_cyy = warp_sum(g_cyy);
                g_op = warp_sum(g_op); g_dd = warp_sum(g_dd);
                g_c0 = warp_sum(g_c0); g_c1 = warp_sum(g_c1); g_c2 = warp_sum(g_c2);
                if (lane == 0) {
                    float *row = acc + (size_t)s_id[j] * ACC_STRIDE;
                    atomicAdd(row + 0, g_mx); atomicAdd(row + 1, g_my);
                    atomicAdd(row + 2, g_cxx); atomicAdd(row + 3, g_cxy); atomicAdd(row + 4, g_cyy);
                    atomicAdd(row + 5, g_op); atomicAdd(row + 6, g_dd);
                    atomicAdd(row + 8, g_c0); atomicAdd(row + 9, g_c1); atomicAdd(row + 10, g_c2);
                }
            }
        }
    }
}

int launch_blend_backward(int P, int W, int H, int64_t R, const uint2 *ranges, const uint32_t *point_list,
                          const GeomPtrs &g, const float *bg, const float *final_T, const uint32_t *n_contrib,
                          const float *dL_dout_color, const float *dL_dout_depth, const float *dL_dout_opacity,
                          int flags, const BlendGradPtrs &o, cudaStream_t s) {
    (void)P;
    if (R <= 0) return 0;
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    const float *dop = (flags & LVDGS_FLAG_OPACITY_GRAD) ? dL_dout_opacity : nullptr;
    LVDGS_PRE(s);
    blend_backward_kernel<<<dim3(gx, gy), BB_THREADS, 0, s>>>(W, H, gx, ranges, point_list, g.means2D, g.conic_opacity,
                                                               g.rgbd, bg, final_T, n_contrib, dL_dout_color,
                                                               dL_dout_depth, dop, o.acc);
    LVDGS_LAUNCHED(s, "blend_backward");
    return 0;
}

}  // namespace lvdgs
