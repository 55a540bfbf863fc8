// blend_forward.cu -- K6: per-16x16-tile front-to-back alpha blending of colour(3) + depth, with the
// opacity (1 - T) and n_touched outputs of the pose-aware fork (reached from
// utils/slam_frontend.py:1493 / utils/slam_backend.py:184 through gaussian_renderer.render).
//
// One CTA per tile, one thread per pixel, each WARP owns an 8x4-pixel block of the tile.  The tile's depth-sorted
// instance list is streamed in batches of 256: each thread gathers one instance's 48 bytes (xy + bounding-box half
// extents, conic + opacity, rgb + depth -- three 128-bit loads from the SoA geometry arrays, which are
// L2-resident) into one shared-memory record and tests the instance's {alpha >= 1/255} bounding box against the eight pixel
// blocks; a ballot per block turns the result into eight 32-bit masks per staging warp.  Each warp then walks only
// the set bits of ITS block's masks, in list order, so an instance that cannot reach a block costs that warp
// nothing (the reference evaluates the exponent for all 256 pixels and discards it).  The test is conservative, so
// every (pixel, Gaussian) pair that passes the reference's two skips is still evaluated and the result is
// unchanged; n_contrib keeps counting list positions.  The per-pair body is branch-free: a pixel that skips the
// Gaussian or terminates on it blends with weight 0, and a finished pixel is "parked" at x = 1e18 (alpha = 0 from then
// on) -- that coordinate doubles as its done flag.
// FP32 FMA/MUFU bound; no tensor cores (there is no dense contraction on this path).
// n_touched is aggregated per warp with a ballot, and only while some pixel of the warp still has T > 0.5
// (T only decreases, so the test T*(1-alpha) > 0.5 can never fire afterwards).
#include "common.cuh"

namespace lvdgs {

constexpr int BF_THREADS = TILE_PIX;
constexpr int BF_WARPS = BF_THREADS / 32;      // 8 warps = 2 x 4 blocks of 8 x 4 pixels
#ifndef LVDGS_BF_SPT
#define LVDGS_BF_SPT 1
#endif
constexpr int BF_SPT = LVDGS_BF_SPT;           // instances staged per thread and barrier pair (see blend_backward.cu)
constexpr int BF_BATCH = BF_THREADS * BF_SPT;

__global__ void __launch_bounds__(BF_THREADS) blend_forward_kernel(int W, int H, int gx, uint32_t capacity, const uint32_t *__restrict__ n_dev, const uint2 *__restrict__ ranges,
                                                                   const uint32_t *__restrict__ point_list,
                                                                   const float4 *__restrict__ means2D,
                                                                   const float4 *__restrict__ conic_opacity,
                                                                   const float4 *__restrict__ rgbd, const uint32_t *__restrict__ tile_order,
                                                                   const float *__restrict__ bg, float *__restrict__ out_color, float *__restrict__ out_depth,
                                                                   float *__restrict__ out_opacity, float *__restrict__ final_T,
                                                                   uint32_t *__restrict__ n_contrib, int32_t *__restrict__ n_touched) {
    __shared__ BlendRec s_rec[BF_BATCH];
    __shared__ uint32_t s_mask[BF_BATCH / 32][BF_WARPS];     // [group of 32 staged entries][pixel block]

    const int tile = tile_order ? (int)__ldg(tile_order + blockIdx.x) : (int)blockIdx.x;   // heaviest tiles first
    const int tile_x = tile % gx, tile_y = tile / gx;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int bx = warp & 1, by = warp >> 1;                       // this warp's 8x4 block inside the tile
    const int px = tile_x * TILE + bx * 8 + (lane & 7);
    const int py = tile_y * TILE + by * 4 + (lane >> 3);
    const bool inside = px < W && py < H;
    float pfx = inside ? (float)px : PIX_PARKED;                   // parked pixels see alpha = 0 for every Gaussian
    const float pfy = (float)py;
    const uint32_t a_rec = smem_u32(s_rec);
    const float tx0 = (float)(tile_x * TILE), ty0 = (float)(tile_y * TILE);   // tile's first pixel centre

    // a speculative launch whose capacity hint was too small has no valid sorted list (the sort retires, see
    // radix_sort.cu); its output is discarded and the tail re-run by the host, so do nothing here
    if (n_dev && __ldg(n_dev) > capacity) return;
    const uint2 range = ranges[tile];
    int todo = (int)(range.y - range.x);
    // a pixel is done (terminated / outside the image) iff it is parked: pfx == PIX_PARKED
    float T = 1.f, C0 = 0.f, C1 = 0.f, C2 = 0.f, D = 0.f;
    uint32_t last_contributor = 0;
    uint32_t batch_first = 0;                                      // list position of the batch's first entry
    int warp_hi = 1;         // some pixel of this warp may still satisfy T(1-alpha) > 0.5

    for (uint32_t base = range.x; todo > 0; base += BF_BATCH, todo -= BF_BATCH, batch_first += BF_BATCH) {
        if (__syncthreads_count(pfx == PIX_PARKED) == BF_THREADS) break;
#pragma unroll
        for (int u = 0; u < BF_SPT; ++u) {
            const int e = u * BF_THREADS + (int)threadIdx.x;       // entry of the batch this thread stages
            uint32_t blocks = 0;                                    // bit (by*2+bx): instance may reach that block
            if (e < todo) {
                const uint32_t id = __ldg(point_list + base + e);
                const float4 m = __ldg(means2D + id);
                const float4 co = __ldg(conic_opacity + id);
                s_rec[e].xy = make_float2(m.x, m.y);
                s_rec[e].id = id;
                // exponent in base 2 with the -1/2 folded in: p2 = A' dx^2 + B' dx dy + C' dy^2, alpha = o 2^p2
                s_rec[e].co = make_float4(-0.5f * LOG2E * co.x, -LOG2E * co.y, -0.5f * LOG2E * co.z, co.w);
                s_rec[e].cd = __ldg(rgbd + id);
                const float rx = m.x - tx0, ry = m.y - ty0;
                uint32_t xb = 0, yb = 0;
                // block column bx spans pixel centres [8bx, 8bx+7]; keep it unless the box [rx-hx, rx+hx] misses it
                if (!(rx + m.z < 0.f) && !(rx - m.z > 7.f)) xb |= 1u;
                if (!(rx + m.z < 8.f) && !(rx - m.z > 15.f)) xb |= 2u;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (!(ry + m.w < 4.f * q) && !(ry - m.w > 4.f * q + 3.f)) yb |= 1u << q;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (yb & (1u << q)) blocks |= xb << (2 * q);
            }
#pragma unroll
            for (int r = 0; r < BF_WARPS; ++r) {
                const uint32_t m = __ballot_sync(0xffffffffu, (blocks >> r) & 1u);
                if (lane == r) s_mask[u * BF_WARPS + warp][r] = m;
            }
        }
        __syncthreads();
        for (int wp = 0; wp < BF_BATCH / 32; ++wp) {
            uint32_t m = s_mask[wp][warp];
            if (__all_sync(0xffffffffu, pfx == PIX_PARKED)) break;
            if (warp_hi) warp_hi = __any_sync(0xffffffffu, pfx != PIX_PARKED && T > 0.5f);
            while (m) {
                const int b = __ffs(m) - 1;
                m &= m - 1;
                const int j = wp * 32 + b;
                // branch-free body: a pixel that skips the Gaussian (alpha < 1/255, power > 0) or terminates on it
                // blends with weight 0 and keeps its state; a terminated pixel is parked and sees alpha = 0 from then on
                const uint32_t a_j = a_rec + (uint32_t)j * (uint32_t)sizeof(BlendRec);
                const float2 xy = lds64(a_j);
                const float4 q = lds128(a_j + 16);
                const float4 cd = lds128(a_j + 32);
                const float dx = xy.x - pfx, dy = xy.y - pfy;
                const float p2 = fmaf(q.z * dy, dy, dx * fmaf(q.x, dx, q.y * dy));
                const float alpha = fminf(0.99f, q.w * ex2_approx(p2));
                const bool ok = p2 <= 0.f && alpha >= 1.f / 255.f;
                const float test_T = T * (1.f - alpha);
                const bool term = ok && test_T < 0.0001f;
                const bool contrib = ok && !term;
                const float wgt = contrib ? alpha * T : 0.f;
                C0 = fmaf(cd.x, wgt, C0); C1 = fmaf(cd.y, wgt, C1); C2 = fmaf(cd.z, wgt, C2); D = fmaf(cd.w, wgt, D);
                T = contrib ? test_T : T;
                last_contributor = contrib ? batch_first + (uint32_t)j + 1u : last_contributor;
                pfx = term ? PIX_PARKED : pfx;
                const bool hit = contrib && test_T > 0.5f;
                if (warp_hi) {
                    const uint32_t bal = __ballot_sync(0xffffffffu, hit);
                    if (bal && lane == 0) atomicAdd(n_touched + lds32(a_j + 8), __popc(bal));
                }
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

int launch_blend_forward(int W, int H, int64_t capacity, const uint32_t *n_dev, const uint2 *ranges, const uint32_t *point_list, const GeomPtrs &g,
                         const uint32_t *tile_order, const float *bg, float *out_color, float *out_depth, float *out_opacity, float *final_T,
                         uint32_t *n_contrib, int32_t *n_touched, cudaStream_t s) {
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    LVDGS_PRE(s);
    blend_forward_kernel<<<gx * gy, BF_THREADS, 0, s>>>(W, H, gx, (uint32_t)min(capacity, (int64_t)0xffffffffll), n_dev, ranges, point_list, g.means2D, g.conic_opacity,
                                                              g.rgbd, tile_order, bg, out_color, out_depth, out_opacity, final_T,
                                                              n_contrib, n_touched);
    LVDGS_LAUNCHED(s, "blend_forward");
    return 0;
}

}  // namespace lvdgs
