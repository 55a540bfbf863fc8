// blend_forward.cu -- K6: per-16x16-tile front-to-back alpha blending of colour(3) + depth, with the
// opacity (1 - T) and n_touched outputs of the pose-aware fork (reached from
// utils/slam_frontend.py:1493 / utils/slam_backend.py:184 through gaussian_renderer.render).
//
// One CTA of 4 warps per tile; a warp owns an 8x8 pixel block, a thread TWO pixels (rows y and y + 4) whose arithmetic
// runs as packed FP32 pairs (FFMA2 / FMUL2 / FADD2, one issue slot for both) -- the same decomposition as the blend
// backward.  The tile's depth-sorted instance list is streamed in batches of 256: each thread gathers two instances'
// 48 bytes (xy + bounding-box half extents, conic + opacity, rgb + depth -- three 128-bit loads from the SoA geometry
// arrays, which are L2-resident) into one shared-memory record each and tests the instance's {alpha >= 1/255} bounding
// box, then the exact ellipse, against the four pixel blocks; a ballot per block turns the result into 32-bit masks.
// Each warp then walks only the set bits of ITS block's masks, in list order, so an instance that cannot reach a block
// costs that warp nothing (the reference evaluates the exponent for all 256 pixels and discards it).  The tests are
// conservative, so every (pixel, Gaussian) pair that passes the reference's two skips is still evaluated and the result
// is unchanged; n_contrib keeps counting list positions.  The per-pair body is branch-free: a pixel that skips the
// Gaussian or terminates on it blends with weight 0, and a finished pixel is "parked" at y = 1e18 (alpha = 0 from then
// on) -- that coordinate doubles as its done flag.
// (One pixel per thread with 8x4 blocks -- finer culling, 1.72 visits of 37 instructions instead of one of 50 -- measured
// 3 % slower: 143.6 vs 139.2 us on the headline view.)
// FP32 FMA/MUFU bound; no tensor cores (there is no dense contraction on this path).
// n_touched is aggregated per warp with a ballot, and only while some pixel of the warp still has T > 0.5
// (T only decreases, so the test T*(1-alpha) > 0.5 can never fire afterwards).
#include "common.cuh"

namespace lvdgs {

#ifndef LVDGS_BF_ELLIPSE
#define LVDGS_BF_ELLIPSE 1      // exact ellipse-vs-block test after the box test (139 vs 141 us)
#endif
constexpr int BF_WARPS = 4, BF_THREADS = BF_WARPS * 32, BF_SPT = 2, BF_BATCH = BF_THREADS * BF_SPT;

#ifndef LVDGS_BF_MINBLOCKS
#define LVDGS_BF_MINBLOCKS 8      // <= 64 registers: 132.6 us against 139 us for 1, 4, 6, 9, 10, 12 (measured)
#endif
__global__ void __launch_bounds__(BF_THREADS, LVDGS_BF_MINBLOCKS) blend_forward_kernel(int W, int H, int gx, uint32_t capacity, const uint32_t *__restrict__ n_dev, const uint2 *__restrict__ ranges,
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
    const int bx = warp & 1, by = warp >> 1;                       // this warp's 8x8 block inside the tile
    const int px = tile_x * TILE + bx * 8 + (lane & 7);
    const int py0 = tile_y * TILE + by * 8 + (lane >> 3), py1 = py0 + 4;
    const bool in0 = px < W && py0 < H, in1 = px < W && py1 < H;
    const float pfx = (float)px;
    // a finished / out-of-image pixel is parked: its -y is -1e18, every Gaussian then evaluates to alpha = 0
    float nya = in0 ? -(float)py0 : -PIX_PARKED, nyb = in1 ? -(float)py1 : -PIX_PARKED;
    const uint32_t a_rec = smem_u32(s_rec);
    const float tx0 = (float)(tile_x * TILE), ty0 = (float)(tile_y * TILE);

    // a speculative launch whose capacity hint was too small has no valid sorted list (the sort retires, see
    // tile_sort.cu); its output is discarded and the tail re-run by the host, so do nothing here
    pdl_wait();                                  // launched behind the tile sort (programmatic dependent launch)
    if (n_dev && __ldg(n_dev) > capacity) return;
    const uint2 range = ranges[tile];
    int todo = (int)(range.y - range.x);
    float Ta = 1.f, Tb = 1.f;
    float C0a = 0.f, C0b = 0.f, C1a = 0.f, C1b = 0.f, C2a = 0.f, C2b = 0.f, Da = 0.f, Db = 0.f;
    uint32_t last0 = 0, last1 = 0;
    uint32_t batch_first = 0;
    int warp_hi = 1;
    auto parked = [&]() { return nya == -PIX_PARKED && nyb == -PIX_PARKED; };

    for (uint32_t base = range.x; todo > 0; base += BF_BATCH, todo -= BF_BATCH, batch_first += BF_BATCH) {
        if (__syncthreads_count(parked()) == BF_THREADS) break;
#pragma unroll
        for (int u = 0; u < BF_SPT; ++u) {
            const int e = u * BF_THREADS + (int)threadIdx.x;
            uint32_t blocks = 0;                                    // bit (by*2+bx): instance may reach that 8x8 block
            if (e < todo) {
                const uint32_t id = __ldg(point_list + base + e);
                const float4 m = __ldg(means2D + id);
                const float4 co = __ldg(conic_opacity + id);
                s_rec[e].xy = make_float2(m.x, m.y);
                s_rec[e].id = id;
                s_rec[e].co = make_float4(-0.5f * LOG2E * co.x, -LOG2E * co.y, -0.5f * LOG2E * co.z, -co.w);   // base-2 conic, MINUS opacity
                s_rec[e].cd = __ldg(rgbd + id);
                const float rx = m.x - tx0, ry = m.y - ty0;
                uint32_t xb = 0;
                if (!(rx + m.z < 0.f) && !(rx - m.z > 7.f)) xb |= 1u;
                if (!(rx + m.z < 8.f) && !(rx - m.z > 15.f)) xb |= 2u;
                if (!(ry + m.w < 0.f) && !(ry - m.w > 7.f)) blocks |= xb;
                if (!(ry + m.w < 8.f) && !(ry - m.w > 15.f)) blocks |= xb << 2;
#if LVDGS_BF_ELLIPSE
                if (blocks) {       // exact ellipse-vs-block test on the survivors of the box test (as in the backward)
                    const float lvl = 2.02f * __logf(255.f * co.w) + 0.02f;
                    const float rA = __fdividef(1.f, co.x), rC = __fdividef(1.f, co.z);
                    uint32_t rest = blocks;
                    while (rest) {
                        const int r = __ffs(rest) - 1;
                        rest &= rest - 1;
                        const float X0 = 8.f * (r & 1), Y0 = 8.f * (r >> 1);
                        if (!ellipse_reaches_rect(rx, ry, co.x, co.y, co.z, rA, rC, lvl, X0, Y0, X0 + 7.f, Y0 + 7.f)) blocks &= ~(1u << r);
                    }
                }
#endif
            }
#pragma unroll
            for (int r = 0; r < BF_WARPS; ++r) {
                const uint32_t mk = __ballot_sync(0xffffffffu, (blocks >> r) & 1u);
                if (lane == r) s_mask[u * BF_WARPS + warp][r] = mk;
            }
        }
        __syncthreads();
        for (int wp = 0; wp < BF_BATCH / 32; ++wp) {
            uint32_t m = s_mask[wp][warp];
            if (__all_sync(0xffffffffu, parked())) break;
            if (warp_hi) warp_hi = __any_sync(0xffffffffu, (nya != -PIX_PARKED && Ta > 0.5f) ||
                                                           (nyb != -PIX_PARKED && Tb > 0.5f));
            while (m) {
                const int b = __ffs(m) - 1;
                m &= m - 1;
                const int j = wp * 32 + b;
                const uint32_t a_j = a_rec + (uint32_t)j * (uint32_t)sizeof(BlendRec);
                const float2 xy = lds64(a_j);
                const float4 q = lds128(a_j + 16);
                const float4 cd = lds128(a_j + 32);
                const float dx = xy.x - pfx;
                const f32x2 dy2 = add2(bc(xy.y), pk(nya, nyb));
                const f32x2 p2 = fma2(mul2(bc(q.z), dy2), dy2, mul2(bc(dx), fma2(bc(q.y), dy2, bc(q.x * dx))));
                const float p2a = lo_of(p2), p2b = hi_of(p2);
                // q.w = -opacity: everything downstream wants -alpha (T - alpha T as ONE fma, weights that accumulate
                // MINUS the colour), so no negation is ever issued
                const float naa = fmaxf(-0.99f, q.w * ex2_approx(p2a)), nab = fmaxf(-0.99f, q.w * ex2_approx(p2b));
                const bool oka = p2a <= 0.f && naa <= -1.f / 255.f, okb = p2b <= 0.f && nab <= -1.f / 255.f;
                const f32x2 T2 = pk(Ta, Tb);
                const f32x2 tT2 = fma2(pk(naa, nab), T2, T2);      // T (1 - alpha)
                const float tTa = lo_of(tT2), tTb = hi_of(tT2);
                const bool ca = oka && !(tTa < 0.0001f), cb = okb && !(tTb < 0.0001f);     // blends; ok && !c = terminates here
                // one select per pixel (-alpha or 0); weight and T follow arithmetically (alpha = 0 leaves T untouched)
                const f32x2 nae2 = pk(ca ? naa : 0.f, cb ? nab : 0.f);
                const f32x2 nw2 = mul2(nae2, T2);
                { const f32x2 c = fma2(bc(cd.x), nw2, pk(C0a, C0b)); C0a = lo_of(c); C0b = hi_of(c); }
                { const f32x2 c = fma2(bc(cd.y), nw2, pk(C1a, C1b)); C1a = lo_of(c); C1b = hi_of(c); }
                { const f32x2 c = fma2(bc(cd.z), nw2, pk(C2a, C2b)); C2a = lo_of(c); C2b = hi_of(c); }
                { const f32x2 c = fma2(bc(cd.w), nw2, pk(Da, Db)); Da = lo_of(c); Db = hi_of(c); }
                { const f32x2 Tn = fma2(nae2, T2, T2); Ta = lo_of(Tn); Tb = hi_of(Tn); }
                const uint32_t idx = batch_first + (uint32_t)j + 1u;
                last0 = ca ? idx : last0; last1 = cb ? idx : last1;
                if (oka) nya = ca ? nya : -PIX_PARKED;
                if (okb) nyb = cb ? nyb : -PIX_PARKED;
                if (warp_hi) {
                    const uint32_t ba = __ballot_sync(0xffffffffu, ca && tTa > 0.5f), bb = __ballot_sync(0xffffffffu, cb && tTb > 0.5f);
                    if ((ba | bb) && lane == 0) atomicAdd(n_touched + lds32(a_j + 8), __popc(ba) + __popc(bb));
                }
            }
        }
    }
    const size_t HW = (size_t)H * W;
    const float bg0 = __ldg(bg + 0), bg1 = __ldg(bg + 1), bg2 = __ldg(bg + 2);
#pragma unroll
    for (int qi = 0; qi < 2; ++qi) {
        if (!(qi ? in1 : in0)) continue;
        const size_t pix = (size_t)(qi ? py1 : py0) * W + px;
        const float T = qi ? Tb : Ta;
        final_T[pix] = T;
        n_contrib[pix] = qi ? last1 : last0;
        out_color[pix] = T * bg0 - (qi ? C0b : C0a);      // the accumulators hold minus the sums
        out_color[HW + pix] = T * bg1 - (qi ? C1b : C1a);
        out_color[2 * HW + pix] = T * bg2 - (qi ? C2b : C2a);
        out_depth[pix] = -(qi ? Db : Da);
        out_opacity[pix] = 1.f - T;
    }
}

int launch_blend_forward(int W, int H, int64_t capacity, const uint32_t *n_dev, const uint2 *ranges, const uint32_t *point_list, const GeomPtrs &g,
                         const uint32_t *tile_order, const float *bg, float *out_color, float *out_depth, float *out_opacity, float *final_T,
                         uint32_t *n_contrib, int32_t *n_touched, cudaStream_t s) {
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    LVDGS_PRE(s);
    LVDGS_CHECK(launch_after_kernel(blend_forward_kernel, dim3(gx * gy), dim3(BF_THREADS), 0, s, W, H, gx, (uint32_t)min(capacity, (int64_t)0xffffffffll),
                                    n_dev, ranges, point_list, g.means2D, g.conic_opacity, g.rgbd, tile_order, bg, out_color, out_depth, out_opacity,
                                    final_T, n_contrib, n_touched));
    LVDGS_LAUNCHED(s, "blend_forward");
    return 0;
}

}  // namespace lvdgs
