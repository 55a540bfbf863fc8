// blend_backward.cu -- K7: back-to-front gradient of the tile blend (reached through
// loss.backward() at utils/slam_frontend.py:1517 and utils/slam_backend.py:306).
//
// The reference issues ten global float atomicAdds per (pixel, Gaussian) pair.  Here:
//  * a CTA of 4 warps owns a 16x16 tile, a warp an 8x8 pixel block, a thread TWO pixels (rows y and y+4), so one
//    warp-level reduction serves 64 pairs, and the two pixels' arithmetic runs as packed FP32 pairs (FFMA2 / FMUL2 /
//    FADD2: one issue slot for both);
//  * instances are staged 256 at a time through shared memory (one 48-byte record each), rearmost first, together
//    with a per-block mask from the same conservative {alpha >= 1/255} bounding-box test as the forward plus an exact
//    ellipse-vs-block test: a warp only visits instances that can reach its block, and the traversal starts at the
//    tile's largest n_contrib instead of the end of the list;
//  * per pixel the recurrences are those of A.4 (T <- T/(1-alpha); what is blended behind the current Gaussian enters
//    only as the scalar S = <B, dL/dpixel>), branch-free: a pixel that does not blend the Gaussian runs the same
//    instructions with alpha = G = 0;
//  * per (warp, Gaussian) the ten partial sums (six geometric moments of m = G dL/dalpha, depth, rgb -- see
//    ACC_STRIDE in common.cuh) are reduced with a transposing butterfly (12 shuffles instead of 5 per value), after
//    which ten lanes each hold one finished sum and commit it with a single warp-wide RED instruction.
#include "common.cuh"

namespace lvdgs {

// Two pixels per thread (rows y and y + 4): a warp owns an 8x8 pixel block, a CTA of 4 warps the tile (1 and 4 pixels per
// thread measured slower: one reduction per 32 pairs / coarser culling).  The two pixels' arithmetic runs as packed
// FP32 pairs (FFMA2 / FMUL2 / FADD2), one issue slot for both.
constexpr int BB_PPT = 2;
constexpr int BB_WARPS = 8 / BB_PPT;
constexpr int BB_THREADS = BB_WARPS * 32;
constexpr int BB_ROWS = 4 * BB_PPT;            // pixel rows per warp block
// instances staged per thread and barrier pair: the four warps of a tile meet once per batch, and the busiest block of a
// batch sets the pace -- a longer batch averages the per-block hit counts out (fewer, better balanced rendezvous)
#ifndef LVDGS_BB_SPT
#define LVDGS_BB_SPT 2
#endif
constexpr int BB_SPT = LVDGS_BB_SPT;
constexpr int BB_BATCH = BB_THREADS * BB_SPT;
#ifndef LVDGS_BB_DIRECT_MAX
#define LVDGS_BB_DIRECT_MAX 2
#endif
constexpr int BB_DIRECT_MAX = LVDGS_BB_DIRECT_MAX;   // up to this many contributing threads: skip the warp reduction
// LVDGS_BB_PIPE: three staging buffers handed between the warps through mbarriers instead of two CTA barriers per batch: a
// warp waits for another one only when that one is more than a batch behind (the per-block hit counts of a batch differ, and
// with __syncthreads the busiest block of every batch sets the pace of all four)
#ifndef LVDGS_BB_PIPE
#define LVDGS_BB_PIPE 0
#endif
constexpr int BB_NBUF = LVDGS_BB_PIPE ? 3 : 1;

// Transposing warp reduction of TEN values per lane: at every butterfly step a lane keeps half of its values and hands
// the other half to its partner, so the value count halves while the lane span halves (5 + 3 + 2 + 1 + 1 = 12 shuffles
// instead of 10 x 5).  Returns the warp-wide sum of value number red10_index(lane); lanes for which that is negative
// hold nothing.  Value i of a lane: i = 5 g + k, group g = lane bit 4, k: see red10_index.
__device__ __forceinline__ int red10_index(int lane) {
    if (lane & 1) return -1;
    const int b3 = (lane >> 3) & 1, b2 = (lane >> 2) & 1, b1 = (lane >> 1) & 1;
    int k;
    if (!b3) k = b2 ? (b1 ? -1 : 2) : b1;          // first three of the five: 0, 1 | 2
    else k = b2 ? -1 : 3 + b1;                     // last two: 3, 4
    return k < 0 ? -1 : 5 * ((lane >> 4) & 1) + k;
}
// lane-constant all-ones / zero masks of lane bits 4, 3, 2, 1: the exchanges select with one LOP3 each, no predicate
struct LaneMasks { uint32_t m16, m8, m4, m2; };
__device__ __forceinline__ LaneMasks lane_masks(int lane) {
    LaneMasks m;
    m.m16 = (lane & 16) ? 0xffffffffu : 0u; m.m8 = (lane & 8) ? 0xffffffffu : 0u;
    m.m4 = (lane & 4) ? 0xffffffffu : 0u; m.m2 = (lane & 2) ? 0xffffffffu : 0u;
    return m;
}
__device__ __forceinline__ float bsel(uint32_t m, float a, float b) {      // m ? a : b
    float d;      // one LOP3 ((a & m) | (b & ~m)); spelled in PTX so that nvcc does not turn it back into predicate + FSEL
    asm("lop3.b32 %0, %1, %2, %3, 0xE4;" : "=f"(d) : "f"(a), "f"(b), "r"(m));
    return d;
}
#define LVDGS_SHX(x, d) __shfl_xor_sync(0xffffffffu, (x), (d))

// shared-memory mbarriers: arrive is non-blocking, waiting is on the phase parity -- the four warps of a tile hand staged
// batches to each other without ever meeting at a CTA-wide barrier (LVDGS_BB_PIPE)
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{ .reg .b64 t; mbarrier.arrive.shared::cta.b64 t, [%0]; }" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{ .reg .pred p;\n"
                 "W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra D;\n"
                 "bra W;\n"
                 "D: }" :: "r"(bar), "r"(parity) : "memory");
}
// returns MINUS the warp-wide sum (the visit's values are negated sums, see the hot loop)
__device__ __forceinline__ float transpose_reduce10(const float (&v)[10], const LaneMasks &L) {
    float w[5], x[3], y[2];
    {   // packed adds: (keep0, keep1) + (recv0, recv1) as one FADD2
        float k[5], r[5];
#pragma unroll
        for (int i = 0; i < 5; ++i) { k[i] = bsel(L.m16, v[i + 5], v[i]); r[i] = LVDGS_SHX(bsel(L.m16, v[i], v[i + 5]), 16); }
        const f32x2 a = add2(pk(k[0], k[1]), pk(r[0], r[1])), b = add2(pk(k[2], k[3]), pk(r[2], r[3]));
        w[0] = lo_of(a); w[1] = hi_of(a); w[2] = lo_of(b); w[3] = hi_of(b); w[4] = k[4] + r[4];
    }
    {   // lower lanes keep w0 w1 w2, upper lanes w3 w4
        const float k0 = bsel(L.m8, w[3], w[0]), k1 = bsel(L.m8, w[4], w[1]), k2 = bsel(L.m8, 0.f, w[2]);
        const float r0 = LVDGS_SHX(bsel(L.m8, w[0], w[3]), 8), r1 = LVDGS_SHX(bsel(L.m8, w[1], w[4]), 8), r2 = LVDGS_SHX(bsel(L.m8, w[2], 0.f), 8);
        const f32x2 a = add2(pk(k0, k1), pk(r0, r1));
        x[0] = lo_of(a); x[1] = hi_of(a); x[2] = k2 + r2;
    }
    {   // lower lanes keep x0 x1, upper lanes x2
        const float k0 = bsel(L.m4, x[2], x[0]), k1 = bsel(L.m4, 0.f, x[1]);
        const float r0 = LVDGS_SHX(bsel(L.m4, x[0], x[2]), 4), r1 = LVDGS_SHX(bsel(L.m4, x[1], 0.f), 4);
        const f32x2 a = add2(pk(k0, k1), pk(r0, r1));
        y[0] = lo_of(a); y[1] = hi_of(a);
    }
    const float z = bsel(L.m2, y[1], y[0]) + LVDGS_SHX(bsel(L.m2, y[0], y[1]), 2);
    return -z - LVDGS_SHX(z, 1);
}

// The same for the six geometric moments alone (3 + 2 + 1 + 1 + 1 = 8 shuffles): value 3 g + k, g = lane bit 4.
__device__ __forceinline__ int red6_index(int lane) {
    if (lane & 3) return -1;
    const int b3 = (lane >> 3) & 1, b2 = (lane >> 2) & 1;
    const int k = b3 ? (b2 ? -1 : 2) : b2;
    return k < 0 ? -1 : 3 * ((lane >> 4) & 1) + k;
}
__device__ __forceinline__ float transpose_reduce6(const float (&v)[10], const LaneMasks &L) {     // also negated
    float w[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) w[i] = bsel(L.m16, v[i + 3], v[i]) + LVDGS_SHX(bsel(L.m16, v[i], v[i + 3]), 16);
    // lower lanes keep w0 w1, upper lanes w2
    const float x0 = bsel(L.m8, w[2], w[0]) + LVDGS_SHX(bsel(L.m8, w[0], w[2]), 8);
    const float x1 = bsel(L.m8, 0.f, w[1]) + LVDGS_SHX(bsel(L.m8, w[1], 0.f), 8);
    float y = bsel(L.m4, x1, x0) + LVDGS_SHX(bsel(L.m4, x0, x1), 4);
    y += LVDGS_SHX(y, 2);
    return -y - LVDGS_SHX(y, 1);
}

// MOMENTS_ONLY: the colour and depth sums are not needed (pose-only backward at SH degree 0 without a depth gradient --
// the tracking loop): six values per (warp, Gaussian) instead of ten.
#ifndef LVDGS_BB_MINBLOCKS
#define LVDGS_BB_MINBLOCKS 8      // <= 64 registers, 8 CTAs per SM: 0.587 against 0.596 ms fwd+bwd on the headline view (measured)
#endif
template <bool MOMENTS_ONLY>
__global__ void __launch_bounds__(BB_THREADS, LVDGS_BB_MINBLOCKS) blend_backward_kernel(
    int W, int H, int gx, const uint2 *__restrict__ ranges, const uint32_t *__restrict__ point_list,
    const float4 *__restrict__ means2D, const float4 *__restrict__ conic_opacity, const float4 *__restrict__ rgbd,
    const uint32_t *__restrict__ tile_order, const float *__restrict__ bg, const float *__restrict__ final_T, const uint32_t *__restrict__ n_contrib,
    const float *__restrict__ dL_dout_color, const float *__restrict__ dL_dout_depth,
    const float *__restrict__ dL_dout_opacity, float *__restrict__ acc, float *__restrict__ zero6) {
    __shared__ BlendRec s_rec[BB_NBUF][BB_BATCH];
    __shared__ uint32_t s_mask[BB_NBUF][BB_BATCH / 32][BB_WARPS];     // [buffer][group of 32 staged entries][pixel block]
    __shared__ uint32_t s_top[BB_WARPS];
#if LVDGS_BB_PIPE
    __shared__ uint64_t s_bar[2 * BB_NBUF];                  // FULL[b] (every thread arrives), EMPTY[b] (one arrival per warp)
#endif

    pdl_wait();      // launched behind the loss kernel when the caller's stream has one right before (programmatic dependent launch)
    // the pose-gradient sum the preprocess backward adds into (it runs after this launch): cleared here, not by a memset
    if (zero6 && blockIdx.x == 0 && threadIdx.x < 6) zero6[threadIdx.x] = 0.f;
    const int tile = tile_order ? (int)__ldg(tile_order + blockIdx.x) : (int)blockIdx.x;   // heaviest tiles first
    const int tile_x = tile % gx, tile_y = tile / gx;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int bx = warp & 1, by = warp >> 1;
    const int px = tile_x * TILE + bx * 8 + (lane & 7);
    const int py0 = tile_y * TILE + by * BB_ROWS + (lane >> 3);
    const float pfx = (float)px;
    const float tx0 = (float)(tile_x * TILE), ty0 = (float)(tile_y * TILE);
    const size_t HW = (size_t)H * W;
    const uint2 range = ranges[tile];
    const uint32_t a_rec0 = smem_u32(s_rec);

    // per-pixel state, packed (lo = row py0, hi = row py0 + 4)
    uint32_t last[2];
    f32x2 npfy2, Tfbgd2, dp0_2, dp1_2, dp2_2, dpd_2; float Ta, Tb, Sa = 0.f, Sb = 0.f;
    {
        float pfy[2], Tf[2], dp0[2], dp1[2], dp2[2], dpd[2], bgd[2];
        const float bg0 = __ldg(bg), bg1 = __ldg(bg + 1), bg2 = __ldg(bg + 2);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int py = py0 + 4 * q;
            const bool inside = px < W && py < H;
            const size_t pix = (size_t)py * W + px;
            pfy[q] = (float)py;
            last[q] = inside ? n_contrib[pix] : 0u;
            Tf[q] = inside ? final_T[pix] : 0.f;
            dp0[q] = inside ? dL_dout_color[pix] : 0.f;
            dp1[q] = inside ? dL_dout_color[HW + pix] : 0.f;
            dp2[q] = inside ? dL_dout_color[2 * HW + pix] : 0.f;
            dpd[q] = (inside && dL_dout_depth) ? dL_dout_depth[pix] : 0.f;
            bgd[q] = bg0 * dp0[q] + bg1 * dp1[q] + bg2 * dp2[q];
            if (inside && dL_dout_opacity) bgd[q] -= dL_dout_opacity[pix];   // d(1 - T_final)/dalpha = +T_final/(1-alpha)
        }
        npfy2 = pk(-pfy[0], -pfy[1]);
        Ta = Tf[0]; Tb = Tf[1];
        Tfbgd2 = pk(Tf[0] * bgd[0], Tf[1] * bgd[1]);
        dp0_2 = pk(dp0[0], dp0[1]); dp1_2 = pk(dp1[0], dp1[1]); dp2_2 = pk(dp2[0], dp2[1]); dpd_2 = pk(dpd[0], dpd[1]);
    }
    // warp-wide and tile-wide max of n_contrib: nothing at or beyond it contributes
    uint32_t wtop = last[0];
#pragma unroll
    for (int q = 1; q < BB_PPT; ++q) wtop = max(wtop, last[q]);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) wtop = max(wtop, __shfl_xor_sync(0xffffffffu, wtop, d));
    if (lane == 0) s_top[warp] = wtop;
    __syncthreads();
    uint32_t top = 0;
#pragma unroll
    for (int k = 0; k < BB_WARPS; ++k) top = max(top, s_top[k]);
    top = min(top, range.y - range.x);

    // value of the transposing reduction this lane ends up holding -> its slot in the accumulator row
    const int red_i = MOMENTS_ONLY ? red6_index(lane) : red10_index(lane);
    const bool commits = red_i >= 0;
    const int slot = red_i < 7 ? red_i : red_i + 1;           // values 0..6 -> slots 0..6, 7..9 -> rgb slots 8..10
    float *const acc_lane = acc + (commits ? slot : 0);
    const LaneMasks LM = lane_masks(lane);

    // entries are visited in decreasing contributor index k = top-1 ... 0
    // entry j of a batch has contributor index k = remaining - 1 - j; "k < n_contrib" is tested as j >= thr
    int thr0 = (int)top - (int)last[0], thr1 = (int)top - (int)last[1];
    // stage(): every thread loads BB_SPT instances of the batch whose first (rearmost) entry has contributor index remaining - 1
    auto stage = [&](int remaining, BlendRec *rec, uint32_t (*msk)[BB_WARPS]) {
        const int nb = min(BB_BATCH, remaining);
#pragma unroll
        for (int u = 0; u < BB_SPT; ++u) {
            const int e = u * BB_THREADS + (int)threadIdx.x;      // entry of the batch this thread stages
            uint32_t blocks = 0;
            if (e < nb) {
                const uint32_t id = __ldg(point_list + range.x + (uint32_t)(remaining - 1 - e));
                const float4 m = __ldg(means2D + id);
                const float4 co = __ldg(conic_opacity + id);
                rec[e].xy = make_float2(m.x, m.y);
                rec[e].id = id;
                rec[e].co = make_float4(-0.5f * LOG2E * co.x, -LOG2E * co.y, -0.5f * LOG2E * co.z, -co.w);   // as forward
                rec[e].cd = __ldg(rgbd + id);
                const float rx = m.x - tx0, ry = m.y - ty0;
                uint32_t xb = 0, yb = 0;
                if (!(rx + m.z < 0.f) && !(rx - m.z > 7.f)) xb |= 1u;
                if (!(rx + m.z < 8.f) && !(rx - m.z > 15.f)) xb |= 2u;
#pragma unroll
                for (int r = 0; r < BB_WARPS / 2; ++r)
                    if (!(ry + m.w < (float)(BB_ROWS * r)) && !(ry - m.w > (float)(BB_ROWS * r + BB_ROWS - 1))) yb |= 1u << r;
#pragma unroll
                for (int r = 0; r < BB_WARPS / 2; ++r)
                    if (yb & (1u << r)) blocks |= xb << (2 * r);
                if (blocks) {       // exact ellipse-vs-block test on the survivors of the box test (as in the forward)
                    const float lvl = 2.02f * __logf(255.f * co.w) + 0.02f;
                    const float rA = __fdividef(1.f, co.x), rC = __fdividef(1.f, co.z);
                    uint32_t rest = blocks;
                    while (rest) {
                        const int r = __ffs(rest) - 1;
                        rest &= rest - 1;
                        const float X0 = 8.f * (r & 1), Y0 = (float)(BB_ROWS * (r >> 1));
                        if (!ellipse_reaches_rect(rx, ry, co.x, co.y, co.z, rA, rC, lvl, X0, Y0, X0 + 7.f, Y0 + (float)(BB_ROWS - 1))) blocks &= ~(1u << r);
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < BB_WARPS; ++r) {
                const uint32_t m = __ballot_sync(0xffffffffu, (blocks >> r) & 1u);
                if (lane == r) msk[u * BB_WARPS + warp][r] = m;
            }
        }
    };
    // process(): this warp visits the staged instances that can reach its block
    auto process = [&](int remaining, uint32_t a_rec, const uint32_t (*msk)[BB_WARPS]) {
        for (int wp = 0; wp < BB_BATCH / 32; ++wp) {
            uint32_t mask = msk[wp][warp];
            {   // entries with contributor index >= wtop (j < remaining - wtop) contribute to no pixel of this warp
                const int first_j = remaining - (int)wtop - wp * 32;
                if (first_j >= 32) mask = 0; else if (first_j > 0) mask &= ~((1u << first_j) - 1u);
            }
            while (mask) {
                const int b = __ffs(mask) - 1;
                mask &= mask - 1;
                const int j = wp * 32 + b;
                const uint32_t a_j = a_rec + (uint32_t)j * (uint32_t)sizeof(BlendRec);
                const float2 xy = lds64(a_j);
                const float4 co = lds128(a_j + 16);
                const float4 cd = lds128(a_j + 32);
                const float dx = xy.x - pfx;
                // phase 1 (cheap): does either pixel blend the Gaussian at all?  The rest of the body is branch-free: a
                // pixel that does not contributes with alpha = 0 and G = 0, which leaves its T / S recurrences untouched
                // (rcp(1) = 1 exactly) and adds zeros to the warp's sums.
                const f32x2 dy2 = add2(bc(xy.y), npfy2);
                const f32x2 p2 = fma2(mul2(bc(co.z), dy2), dy2, mul2(bc(dx), fma2(bc(co.y), dy2, bc(co.x * dx))));
                const float p2a = lo_of(p2), p2b = hi_of(p2);
                const float ga = ex2_approx(p2a), gb = ex2_approx(p2b);
                // co.w = -opacity; the recurrences below are written in -alpha, and every sum of this visit comes out
                // NEGATED (m = G dL/dalpha and the blend weight both carry the sign), which the commit undoes for free
                const float naa = fmaxf(-0.99f, co.w * ga), nab = fmaxf(-0.99f, co.w * gb);
                const bool oka = j >= thr0 && p2a <= 0.f && naa <= -1.f / 255.f;
                const bool okb = j >= thr1 && p2b <= 0.f && nab <= -1.f / 255.f;
                const bool valid = oka || okb;
                const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
                if (!vmask) continue;
                const f32x2 G2 = pk(oka ? ga : 0.f, okb ? gb : 0.f);
                const f32x2 nal2 = pk(oka ? naa : 0.f, okb ? nab : 0.f);
                const f32x2 one_m2 = add2(nal2, bc(1.f));
                const f32x2 inv2 = pk(rcp_approx(lo_of(one_m2)), rcp_approx(hi_of(one_m2)));   // 1 - alpha >= 0.01
                const f32x2 T2 = mul2(pk(Ta, Tb), inv2); Ta = lo_of(T2); Tb = hi_of(T2);
                const f32x2 S2o = pk(Sa, Sb);
                // the colour / depth blended BEHIND this Gaussian enters only through its dot product with dL/dpixel:
                // S = <B, dp> obeys the same recurrence as B itself (S <- S + alpha (<c, dp> - S))
                const f32x2 cdp2 = fma2(bc(cd.w), dpd_2, fma2(bc(cd.z), dp2_2, fma2(bc(cd.y), dp1_2, mul2(bc(cd.x), dp0_2))));
                const f32x2 ne2 = fma2(cdp2, bc(-1.f), S2o);       // S - <c, dp>
                f32x2 ndL2 = mul2(ne2, T2);                       // -dL/dalpha, first term
                ndL2 = fma2(inv2, Tfbgd2, ndL2);                  // + T_final / (1 - alpha) * <bg, dp>
                const f32x2 S2n = fma2(nal2, ne2, S2o); Sa = lo_of(S2n); Sb = hi_of(S2n);
                const f32x2 m2 = mul2(G2, ndL2);                  // -m
                const f32x2 mdx2 = mul2(m2, bc(dx)), mdy2 = mul2(m2, dy2);
                float v[10];      // NEGATED accumulator-row slots 0..6, 8..10 (ACC_STRIDE layout in common.cuh)
                v[0] = hsum(mdx2); v[1] = hsum(mdy2);
                v[2] = dx * v[0];                                 // both pixels share dx
                v[3] = hsum(mul2(mdx2, dy2)); v[4] = hsum(mul2(mdy2, dy2));
                v[5] = hsum(m2);
                if (!MOMENTS_ONLY) {
                    const f32x2 wgt2 = mul2(nal2, T2);            // -alpha T
                    v[6] = hsum(mul2(wgt2, dpd_2));
                    v[7] = hsum(mul2(wgt2, dp0_2)); v[8] = hsum(mul2(wgt2, dp1_2)); v[9] = hsum(mul2(wgt2, dp2_2));
                } else {
                    v[6] = v[7] = v[8] = v[9] = 0.f;
                }
                if (vmask) {
                    const uint32_t id = lds32(a_j + 8);
                    if (__popc(vmask) <= BB_DIRECT_MAX) {
                        // a Gaussian's edge often reaches only one or two threads of the block: their partial sums go
                        // straight to the accumulator row (10 REDs) instead of through the 60-instruction reduction
                        if (valid) {
                            float *row = acc + (size_t)id * ACC_STRIDE;
                            atomicAdd(row + 0, -v[0]); atomicAdd(row + 1, -v[1]); atomicAdd(row + 2, -v[2]); atomicAdd(row + 3, -v[3]);
                            atomicAdd(row + 4, -v[4]); atomicAdd(row + 5, -v[5]);
                            if (!MOMENTS_ONLY) {
                                atomicAdd(row + 6, -v[6]);
                                atomicAdd(row + 8, -v[7]); atomicAdd(row + 9, -v[8]); atomicAdd(row + 10, -v[9]);
                            }
                        }
                    } else {
                        const float sum = MOMENTS_ONLY ? transpose_reduce6(v, LM) : transpose_reduce10(v, LM);
                        if (commits) atomicAdd(acc_lane + (size_t)id * ACC_STRIDE, sum);
                    }
                }
            }
        }
    };
#if !LVDGS_BB_PIPE
    for (int remaining = (int)top; remaining > 0; remaining -= BB_BATCH, thr0 -= BB_BATCH, thr1 -= BB_BATCH) {
        __syncthreads();
        stage(remaining, s_rec[0], s_mask[0]);
        __syncthreads();
        process(remaining, a_rec0, s_mask[0]);
    }
#else
    const int nbatch = ((int)top + BB_BATCH - 1) / BB_BATCH;
    const uint32_t a_bar = smem_u32(s_bar);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int b = 0; b < BB_NBUF; ++b) { mbar_init(a_bar + 8 * b, BB_THREADS); mbar_init(a_bar + 8 * (BB_NBUF + b), BB_WARPS); }
    }
    __syncthreads();
    for (int k = 0; k < 2 && k < nbatch; ++k) {             // prologue: two batches in flight
        stage((int)top - k * BB_BATCH, s_rec[k], s_mask[k]);
        mbar_arrive(a_bar + 8 * k);
    }
    int b = 0, ph = 0;                                       // buffer and phase parity of batch k
    for (int k = 0, remaining = (int)top; k < nbatch; ++k, remaining -= BB_BATCH, thr0 -= BB_BATCH, thr1 -= BB_BATCH) {
        mbar_wait(a_bar + 8 * b, (uint32_t)ph);              // every thread's share of batch k has been staged
        process(remaining, a_rec0 + (uint32_t)b * (uint32_t)(BB_BATCH * sizeof(BlendRec)), s_mask[b]);
        __syncwarp();
        if (lane == 0) mbar_arrive(a_bar + 8 * (BB_NBUF + b));      // this warp is done with buffer b
        if (k + 2 < nbatch) {
            // batch k + 2 goes where batch k - 1 was: wait until all four warps have left that one (they are at most a
            // batch behind in the common case, so this rarely blocks)
            const int b2 = b == 0 ? 2 : b - 1;
            if (k >= 1) mbar_wait(a_bar + 8 * (BB_NBUF + b2), (uint32_t)(b == 0 ? ph ^ 1 : ph));
            stage(remaining - 2 * BB_BATCH, s_rec[b2], s_mask[b2]);
            mbar_arrive(a_bar + 8 * b2);
        }
        if (++b == BB_NBUF) { b = 0; ph ^= 1; }
    }
#endif
}

int launch_blend_backward(int P, int W, int H, int64_t R, const uint2 *ranges, const uint32_t *point_list,
                          const uint32_t *tile_order, const GeomPtrs &g, const float *bg, const float *final_T, const uint32_t *n_contrib,
                          const float *dL_dout_color, const float *dL_dout_depth, const float *dL_dout_opacity,
                          int flags, bool moments_only, const BlendGradPtrs &o, float *zero6, cudaStream_t s) {
    (void)P;
    if (R <= 0) return 0;
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    const float *dop = (flags & LVDGS_FLAG_OPACITY_GRAD) ? dL_dout_opacity : nullptr;
    LVDGS_PRE(s);
    if (moments_only)
        LVDGS_CHECK(launch_after_kernel(blend_backward_kernel<true>, dim3(gx * gy), dim3(BB_THREADS), 0, s, W, H, gx, ranges, point_list, g.means2D,
                                        g.conic_opacity, g.rgbd, tile_order, bg, final_T, n_contrib, dL_dout_color, nullptr, dop, o.acc, zero6));
    else
        LVDGS_CHECK(launch_after_kernel(blend_backward_kernel<false>, dim3(gx * gy), dim3(BB_THREADS), 0, s, W, H, gx, ranges, point_list, g.means2D,
                                        g.conic_opacity, g.rgbd, tile_order, bg, final_T, n_contrib, dL_dout_color, dL_dout_depth, dop, o.acc, zero6));
    LVDGS_LAUNCHED(s, "blend_backward");
    return 0;
}

}  // namespace lvdgs
