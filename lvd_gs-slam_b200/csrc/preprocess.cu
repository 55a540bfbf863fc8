// preprocess.cu -- K1 preprocess forward (with the first half of the K2 tile-count scan fused in), binning_count /
// binning_prep (K2 + K5: per-tile counts -> tile ranges, key offsets, R; digit histograms for the onesweep path),
// K3 key emission (into per-tile segments, or in emission order for the onesweep path), K10 markVisible.
//
// Behaviour follows SURVEY.md Appendix A.1 / A.2 (the reference's `preprocessCUDA`, `duplicateWithKeys`,
// `identifyTileRanges`, `checkFrustum`, reached from utils/slam_frontend.py:1493 and utils/slam_backend.py:184
// through the missing gaussian_renderer shim).  Design is B200-first:
//  * the [P,3] AoS inputs are staged through shared memory with fully coalesced loads, quaternions are
//    one 128-bit load, all intermediate state is SoA with 8/16-byte vector stores;
//  * every operation on the chain to an integer output is an explicit IEEE intrinsic (canonical arithmetic);
//  * key emission is block-cooperative: a block owns 256 consecutive Gaussians, and its threads walk the block's
//    contiguous span of instances (binary search in shared memory) instead of one thread looping over its own tile
//    rectangle; an instance either claims a slot of its tile's segment (tile_sort.cu) or, for the global onesweep
//    sort, is stored at its emission position with coalesced stores.
#include "common.cuh"

namespace lvdgs {

constexpr int PRE_THREADS = 256;

__device__ __constant__ float SH_C0 = 0.28209479177387814f;
__device__ __constant__ float SH_C1 = 0.4886025119029199f;
__device__ __constant__ float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                          -1.0925484305920792f, 0.5462742152960396f};
__device__ __constant__ float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                          0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                                          -0.5900435899266435f};

// coalesced load of 256 x 3 floats into shared memory (block-strided), then each thread picks its triple
__device__ __forceinline__ float3 load3_staged(const float *__restrict__ base, int P, float *smem, int bid) {
    const int blk0 = bid * PRE_THREADS;
    const int n = min(PRE_THREADS, P - blk0) * 3;
    const float *src = base + (size_t)blk0 * 3;
    for (int k = threadIdx.x; k < n; k += PRE_THREADS) smem[k] = __ldg(src + k);
    __syncthreads();
    float3 v = make_float3(smem[3 * threadIdx.x], smem[3 * threadIdx.x + 1], smem[3 * threadIdx.x + 2]);
    __syncthreads();
    return v;
}

__device__ __forceinline__ void quat_to_R(float4 q, float R[9]) {
    const float r = q.x, x = q.y, y = q.z, z = q.w;
    R[0] = __fmaf_rn(-2.f, __fmaf_rn(z, z, __fmul_rn(y, y)), 1.f);
    R[1] = __fmul_rn(2.f, __fmaf_rn(x, y, -__fmul_rn(r, z)));
    R[2] = __fmul_rn(2.f, __fmaf_rn(x, z, __fmul_rn(r, y)));
    R[3] = __fmul_rn(2.f, __fmaf_rn(x, y, __fmul_rn(r, z)));
    R[4] = __fmaf_rn(-2.f, __fmaf_rn(z, z, __fmul_rn(x, x)), 1.f);
    R[5] = __fmul_rn(2.f, __fmaf_rn(y, z, -__fmul_rn(r, x)));
    R[6] = __fmul_rn(2.f, __fmaf_rn(x, z, -__fmul_rn(r, y)));
    R[7] = __fmul_rn(2.f, __fmaf_rn(y, z, __fmul_rn(r, x)));
    R[8] = __fmaf_rn(-2.f, __fmaf_rn(y, y, __fmul_rn(x, x)), 1.f);
}

__device__ __forceinline__ void cov3d_from_scale_rot(float3 s, float mod, float4 q, float c[6]) {
    float R[9];
    quat_to_R(q, R);
    const float s0 = __fmul_rn(mod, s.x), s1 = __fmul_rn(mod, s.y), s2 = __fmul_rn(mod, s.z);
    float A[9];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        A[a * 3 + 0] = __fmul_rn(R[a * 3 + 0], s0);
        A[a * 3 + 1] = __fmul_rn(R[a * 3 + 1], s1);
        A[a * 3 + 2] = __fmul_rn(R[a * 3 + 2], s2);
    }
    c[0] = dot3c(A[0], A[0], A[1], A[1], A[2], A[2]);
    c[1] = dot3c(A[0], A[3], A[1], A[4], A[2], A[5]);
    c[2] = dot3c(A[0], A[6], A[1], A[7], A[2], A[8]);
    c[3] = dot3c(A[3], A[3], A[4], A[4], A[5], A[5]);
    c[4] = dot3c(A[3], A[6], A[4], A[7], A[5], A[8]);
    c[5] = dot3c(A[6], A[6], A[7], A[7], A[8], A[8]);
}

// EWA projection (A.1 step 4): returns the dilated 2D covariance (a,b,c)
__device__ __forceinline__ float3 ewa_cov2d(float tx, float ty, float tz, float fx, float fy, float tanfovx,
                                            float tanfovy, const float c3[6], const float *view) {
    const float limx = __fmul_rn(1.3f, tanfovx), limy = __fmul_rn(1.3f, tanfovy);
    const float txtz = __fdiv_rn(tx, tz), tytz = __fdiv_rn(ty, tz);
    const float txc = __fmul_rn(fminf(limx, fmaxf(-limx, txtz)), tz);
    const float tyc = __fmul_rn(fminf(limy, fmaxf(-limy, tytz)), tz);
    const float tz2 = __fmul_rn(tz, tz);
    const float J00 = __fdiv_rn(fx, tz);
    const float J02 = __fdiv_rn(-__fmul_rn(fx, txc), tz2);
    const float J11 = __fdiv_rn(fy, tz);
    const float J12 = __fdiv_rn(-__fmul_rn(fy, tyc), tz2);
    float m0[3], m1[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float R0 = view[4 * k + 0], R1 = view[4 * k + 1], R2 = view[4 * k + 2];
        m0[k] = __fmaf_rn(J02, R2, __fmul_rn(J00, R0));
        m1[k] = __fmaf_rn(J12, R2, __fmul_rn(J11, R1));
    }
    float u0[3], u1[3];
    u0[0] = dot3c(c3[0], m0[0], c3[1], m0[1], c3[2], m0[2]);
    u0[1] = dot3c(c3[1], m0[0], c3[3], m0[1], c3[4], m0[2]);
    u0[2] = dot3c(c3[2], m0[0], c3[4], m0[1], c3[5], m0[2]);
    u1[0] = dot3c(c3[0], m1[0], c3[1], m1[1], c3[2], m1[2]);
    u1[1] = dot3c(c3[1], m1[0], c3[3], m1[1], c3[4], m1[2]);
    u1[2] = dot3c(c3[2], m1[0], c3[4], m1[1], c3[5], m1[2]);
    float3 cov;
    cov.x = __fadd_rn(dot3c(m0[0], u0[0], m0[1], u0[1], m0[2], u0[2]), 0.3f);
    cov.y = dot3c(m0[0], u1[0], m0[1], u1[1], m0[2], u1[2]);
    cov.z = __fadd_rn(dot3c(m1[0], u1[0], m1[1], u1[1], m1[2], u1[2]), 0.3f);
    return cov;
}

__device__ __forceinline__ float3 sh_to_rgb(int deg, int M, const float *__restrict__ sh, float3 p, const float *campos,
                                            uint8_t &clamped) {
    float3 res;
    float r[3];
    if (deg == 0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) r[c] = __fmaf_rn(SH_C0, __ldg(sh + c), 0.5f);
    } else {
        float3 d = make_float3(__fsub_rn(p.x, campos[0]), __fsub_rn(p.y, campos[1]), __fsub_rn(p.z, campos[2]));
        const float inv = __fdiv_rn(1.f, __fsqrt_rn(dot3c(d.x, d.x, d.y, d.y, d.z, d.z)));
        const float x = d.x * inv, y = d.y * inv, z = d.z * inv;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
#define SHC(k) __ldg(sh + (k) * 3 + c)
            float v = SH_C0 * SHC(0);
            v = v - SH_C1 * y * SHC(1) + SH_C1 * z * SHC(2) - SH_C1 * x * SHC(3);
            if (deg > 1) {
                const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                v = v + SH_C2[0] * xy * SHC(4) + SH_C2[1] * yz * SHC(5) + SH_C2[2] * (2.f * zz - xx - yy) * SHC(6) +
                    SH_C2[3] * xz * SHC(7) + SH_C2[4] * (xx - yy) * SHC(8);
                if (deg > 2) {
                    v = v + SH_C3[0] * y * (3.f * xx - yy) * SHC(9) + SH_C3[1] * xy * z * SHC(10) +
                        SH_C3[2] * y * (4.f * zz - xx - yy) * SHC(11) +
                        SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy) * SHC(12) +
                        SH_C3[4] * x * (4.f * zz - xx - yy) * SHC(13) + SH_C3[5] * z * (xx - yy) * SHC(14) +
                        SH_C3[6] * x * (xx - 3.f * yy) * SHC(15);
                }
            }
#undef SHC
            r[c] = v + 0.5f;
        }
    }
    clamped = 0;
#pragma unroll
    for (int c = 0; c < 3; ++c)
        if (r[c] < 0.f) { clamped |= (uint8_t)(1u << c); r[c] = 0.f; }
    res = make_float3(r[0], r[1], r[2]);
    return res;
}

struct PreArgs {
    int P, D, M, W, H, gx, gy;
    float tanfovx, tanfovy, fx, fy, mod;
    const float *means3D, *colors_precomp, *opacities, *scales, *rotations, *cov3D_precomp, *view, *proj, *shs, *campos;
    int32_t *radii;
    int32_t *n_touched;        // zeroed here (the blend forward counts into it)
    uint32_t *zero_base;       // tile difference grid + digit histograms + tile cursors: cleared here, every block its share
    int zero_words;            // (binning_count, the first kernel to add into them, runs after this launch has finished)
    GeomPtrs g;
};

#ifndef LVDGS_PF_MINBLOCKS
#define LVDGS_PF_MINBLOCKS 1
#endif
__global__ void __launch_bounds__(PRE_THREADS, LVDGS_PF_MINBLOCKS) preprocess_forward_kernel(const PreArgs a) {
    __shared__ float stage[PRE_THREADS * 9];
    __shared__ CameraConst cam;
    __shared__ uint32_t s_warp[PRE_THREADS / 32];
    const int bid = blockIdx.x;
    pdl_wait();                // may be launched behind the previous kernel of the stream (tracking: the pose step)
    {
        const int per_block = (a.zero_words + (int)gridDim.x - 1) / (int)gridDim.x;
        const int lo = bid * per_block, hi = min(a.zero_words, lo + per_block);
        for (int k = lo + (int)threadIdx.x; k < hi; k += PRE_THREADS) a.zero_base[k] = 0u;
    }
    if (threadIdx.x < 16) cam.view[threadIdx.x] = __ldg(a.view + threadIdx.x);
    else if (threadIdx.x < 32) cam.proj[threadIdx.x - 16] = __ldg(a.proj + threadIdx.x - 16);
    else if (threadIdx.x < 35) cam.campos[threadIdx.x - 32] = __ldg(a.campos + threadIdx.x - 32);
    const int i = bid * PRE_THREADS + threadIdx.x;
    // all three [P,3] inputs are staged with ONE barrier pair: the loads of means / scales / colour are in flight together
    const bool staged_color = a.colors_precomp != nullptr || a.M == 1;
    const float *color_src = a.colors_precomp ? a.colors_precomp : a.shs;
    {
        const int blk0 = bid * PRE_THREADS;
        const int n = min(PRE_THREADS, a.P - blk0) * 3;
        const size_t off = (size_t)blk0 * 3;
        for (int k = threadIdx.x; k < n; k += PRE_THREADS) {
            stage[k] = __ldg(a.means3D + off + k);
            if (a.scales) stage[PRE_THREADS * 3 + k] = __ldg(a.scales + off + k);
            if (staged_color) stage[PRE_THREADS * 6 + k] = __ldg(color_src + off + k);
        }
    }
    // the per-Gaussian 128-bit / 32-bit inputs are requested before the barrier as well, culled or not: one memory round
    // trip for the whole kernel instead of three dependent ones (20 B more per culled Gaussian, all coalesced)
    float4 q_in = make_float4(1.f, 0.f, 0.f, 0.f);
    float opac_in = 0.f;
    if (i < a.P) {
        if (!a.cov3D_precomp) q_in = __ldg(reinterpret_cast<const float4 *>(a.rotations) + i);
        opac_in = __ldg(a.opacities + i);
    }
    __syncthreads();                                            // also publishes `cam`
    const float3 p = make_float3(stage[3 * threadIdx.x], stage[3 * threadIdx.x + 1], stage[3 * threadIdx.x + 2]);
    float3 sc = make_float3(0.f, 0.f, 0.f), sh0 = make_float3(0.f, 0.f, 0.f);
    if (a.scales) sc = make_float3(stage[PRE_THREADS * 3 + 3 * threadIdx.x], stage[PRE_THREADS * 3 + 3 * threadIdx.x + 1], stage[PRE_THREADS * 3 + 3 * threadIdx.x + 2]);
    if (staged_color) sh0 = make_float3(stage[PRE_THREADS * 6 + 3 * threadIdx.x], stage[PRE_THREADS * 6 + 3 * threadIdx.x + 1], stage[PRE_THREADS * 6 + 3 * threadIdx.x + 2]);
    const bool in_range = i < a.P;

    int32_t radius = 0;
    uint32_t touched = 0;
    float depth = 0.f;
    float4 pix = make_float4(0.f, 0.f, -1e30f, -1e30f);
    float4 con_o = make_float4(0.f, 0.f, 0.f, 0.f);
    float3 rgb = make_float3(0.f, 0.f, 0.f);
    short4 rect = make_short4(0, 0, 0, 0);
    uint8_t clamped = 0;

    const float *V = cam.view, *Pj = cam.proj;
    const float tx = __fadd_rn(dot3c(V[0], p.x, V[4], p.y, V[8], p.z), V[12]);
    const float ty = __fadd_rn(dot3c(V[1], p.x, V[5], p.y, V[9], p.z), V[13]);
    const float tz = __fadd_rn(dot3c(V[2], p.x, V[6], p.y, V[10], p.z), V[14]);
    if (in_range && tz > 0.2f) {
        const float hx = __fadd_rn(dot3c(Pj[0], p.x, Pj[4], p.y, Pj[8], p.z), Pj[12]);
        const float hy = __fadd_rn(dot3c(Pj[1], p.x, Pj[5], p.y, Pj[9], p.z), Pj[13]);
        const float hw = __fadd_rn(dot3c(Pj[3], p.x, Pj[7], p.y, Pj[11], p.z), Pj[15]);
        const float pw = __fdiv_rn(1.f, __fadd_rn(hw, 0.0000001f));
        const float px = __fmul_rn(hx, pw), py = __fmul_rn(hy, pw);
        float c3[6];
        if (a.cov3D_precomp) {
#pragma unroll
            for (int k = 0; k < 6; ++k) c3[k] = __ldg(a.cov3D_precomp + (size_t)i * 6 + k);
        } else {
            cov3d_from_scale_rot(sc, a.mod, q_in, c3);
        }
        const float3 cov = ewa_cov2d(tx, ty, tz, a.fx, a.fy, a.tanfovx, a.tanfovy, c3, V);
        const float det = __fmaf_rn(cov.x, cov.z, -__fmul_rn(cov.y, cov.y));
        if (det != 0.f) {
            const float det_inv = __fdiv_rn(1.f, det);
            const float mid = __fmul_rn(0.5f, __fadd_rn(cov.x, cov.z));
            const float sq = __fsqrt_rn(fmaxf(0.1f, __fmaf_rn(mid, mid, -det)));
            const float lam = fmaxf(__fadd_rn(mid, sq), __fsub_rn(mid, sq));
            const float rad = ceilf(__fmul_rn(3.f, __fsqrt_rn(lam)));
            const float pix_x = __fmul_rn(__fmaf_rn(__fadd_rn(px, 1.f), (float)a.W, -1.f), 0.5f);
            const float pix_y = __fmul_rn(__fmaf_rn(__fadd_rn(py, 1.f), (float)a.H, -1.f), 0.5f);
            const int irad = (int)rad;
            const float fr = (float)irad;
            const int rminx = min(a.gx, max(0, (int)__fdiv_rn(__fsub_rn(pix_x, fr), 16.f)));
            const int rminy = min(a.gy, max(0, (int)__fdiv_rn(__fsub_rn(pix_y, fr), 16.f)));
            const int rmaxx = min(a.gx, max(0, (int)__fdiv_rn(__fadd_rn(__fadd_rn(pix_x, fr), 15.f), 16.f)));
            const int rmaxy = min(a.gy, max(0, (int)__fdiv_rn(__fadd_rn(__fadd_rn(pix_y, fr), 15.f), 16.f)));
            const int area = (rmaxx - rminx) * (rmaxy - rminy);
            if (area != 0) {
                if (a.colors_precomp) rgb = sh0;
                else if (a.M == 1 && a.D == 0) {
                    float r[3] = {__fmaf_rn(SH_C0, sh0.x, 0.5f), __fmaf_rn(SH_C0, sh0.y, 0.5f), __fmaf_rn(SH_C0, sh0.z, 0.5f)};
#pragma unroll
                    for (int c = 0; c < 3; ++c)
                        if (r[c] < 0.f) { clamped |= (uint8_t)(1u << c); r[c] = 0.f; }
                    rgb = make_float3(r[0], r[1], r[2]);
                } else {
                    rgb = sh_to_rgb(a.D, a.M, a.shs + (size_t)i * a.M * 3, p, cam.campos, clamped);
                }
                depth = tz;
                radius = irad;
                const float opac = opac_in;
                // Half extents of the axis-aligned box around {alpha >= 1/255} = {d^T conic d <= 2 ln(255 o)}: the blend
                // kernels use it to drop (pixel block, Gaussian) pairs that upstream would evaluate and then skip.
                // Conservative (1% + 0.02 slack on the level, so float noise in `power` can never flip a decision);
                // o <= 1/255 gives a negative level -> extents stay -1e30 and the Gaussian is never evaluated.
                const float lvl = 2.02f * __logf(255.f * opac) + 0.02f;
                float hx = -1e30f, hy = -1e30f;
                if (lvl > 0.f) { hx = sqrtf(lvl * cov.x) * 1.0001f + 0.01f; hy = sqrtf(lvl * cov.z) * 1.0001f + 0.01f; }
                if (!(opac == opac) || !(lvl == lvl)) { hx = 1e30f; hy = 1e30f; }   // NaN inputs: never cull
                pix = make_float4(pix_x, pix_y, hx, hy);
                con_o = make_float4(__fmul_rn(cov.z, det_inv), __fmul_rn(-cov.y, det_inv), __fmul_rn(cov.x, det_inv), opac);
                rect = make_short4((short)rminx, (short)rminy, (short)rmaxx, (short)rmaxy);
                touched = (uint32_t)area;
            }
        }
    }
    if (in_range) {
        a.radii[i] = radius;
        a.g.depths[i] = depth;
        a.g.means2D[i] = pix;
        a.g.conic_opacity[i] = con_o;
        a.g.rgbd[i] = make_float4(rgb.x, rgb.y, rgb.z, depth);
        a.g.rect[i] = rect;
        a.g.tiles_touched[i] = touched;
        a.g.clamped[i] = clamped;
        a.n_touched[i] = 0;
    }
    // ---- K2, first half: this block's instance count (binning_prep scans the block sums, emit_keys scans inside) ----
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t sum = touched;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
    if (lane == 0) s_warp[warp] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
#pragma unroll
        for (int k = 0; k < PRE_THREADS / 32; ++k) t += s_warp[k];
        a.g.block_sums[bid] = t;
    }
}

int launch_preprocess_forward(const lvdgs_raster_params &p, const float *means3D, const float *colors_precomp,
                              const float *opacities, const float *scales, const float *rotations,
                              const float *cov3D_precomp, const float *view, const float *proj, const float *shs,
                              const float *campos, int32_t *radii, int32_t *n_touched, const GeomPtrs &g, const ImgPtrs &im, cudaStream_t s) {
    PreArgs a;
    {
        lvdgs_img_layout il;
        lvdgs_get_img_layout(p.width, p.height, &il);
        a.zero_base = reinterpret_cast<uint32_t *>(im.tile_grid);
        a.zero_words = (int)((il.total - il.tile_grid) / sizeof(uint32_t));
    }
    a.n_touched = n_touched;
    a.P = p.P; a.D = p.sh_degree; a.M = p.sh_coeffs; a.W = p.width; a.H = p.height;
    a.gx = (p.width + TILE - 1) / TILE; a.gy = (p.height + TILE - 1) / TILE;
    a.tanfovx = p.tan_fovx; a.tanfovy = p.tan_fovy;
    a.fx = (float)p.width / (2.f * p.tan_fovx); a.fy = (float)p.height / (2.f * p.tan_fovy);
    a.mod = p.scale_modifier;
    a.means3D = means3D; a.colors_precomp = colors_precomp; a.opacities = opacities; a.scales = scales;
    a.rotations = rotations; a.cov3D_precomp = cov3D_precomp; a.view = view; a.proj = proj; a.shs = shs;
    a.campos = campos; a.radii = radii; a.g = g;
    LVDGS_PRE(s);
    LVDGS_CHECK(launch_after_kernel(preprocess_forward_kernel, dim3(ceil_div(p.P, PRE_THREADS)), dim3(PRE_THREADS), 0, s, a));
    LVDGS_LAUNCHED(s, "preprocess_forward");
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// binning_count: per-tile instance counts (and, for the onesweep path only, the histograms of the four depth digits
// of the sort keys) WITHOUT touching the instances: every visible Gaussian contributes its tile rect as a 2-D
// difference (4 corner updates instead of one update per covered tile) and `tiles_touched` to the bin of each depth
// byte.  A few large CTAs with shared-memory-privatised tables, flushed once, so global atomics are ~1 per table
// cell per CTA.
// binning_prep (below): one CTA.  Scans the preprocess blocks' instance counts (-> per-block key offsets, R),
// integrates the difference array into per-tile instance counts, turns them into the tile ranges (K5 -- no pass over
// the sorted keys is needed: range(t) = [sum of counts before t, + count(t))), records the longest list, orders the
// tiles by decreasing list length, and -- onesweep path -- derives and exclusive-scans the digit histograms.
// ---------------------------------------------------------------------------------------------------------
constexpr int COUNT_THREADS = 1024;
constexpr int PREP_GRID_SMEM = 10240;     // difference-array cells kept in shared memory (covers 1920x1080: 121 x 69)

__global__ void __launch_bounds__(COUNT_THREADS) binning_count_kernel(int P, int gx, int gy, const uint32_t *__restrict__ tiles_touched,
                                                                      const short4 *__restrict__ rects, const float *__restrict__ depths,
                                                                      int32_t *__restrict__ grid_g, uint32_t *__restrict__ hist_g) {
    extern __shared__ int32_t s_tab[];          // [4*256] depth-digit histograms, then the grid if it fits
    const int gw = gx + 1, cells = gw * (gy + 1);
    const bool in_smem = cells <= PREP_GRID_SMEM;
    uint32_t *s_hist = reinterpret_cast<uint32_t *>(s_tab);
    int32_t *s_grid = s_tab + 4 * SORT_BINS;
    for (int k = threadIdx.x; k < 4 * SORT_BINS + (in_smem ? cells : 0); k += COUNT_THREADS) s_tab[k] = 0;
    __syncthreads();
    pdl_wait();                                  // launched behind preprocess_forward (programmatic dependent launch)
    int32_t *grid = in_smem ? s_grid : grid_g;
    for (int i = blockIdx.x * COUNT_THREADS + threadIdx.x; i < P; i += gridDim.x * COUNT_THREADS) {
        const uint32_t touched = __ldg(tiles_touched + i);
        if (!touched) continue;
        const short4 rc = __ldg(rects + i);
        atomicAdd(grid + rc.y * gw + rc.x, 1);
        atomicAdd(grid + rc.y * gw + rc.z, -1);
        atomicAdd(grid + rc.w * gw + rc.x, -1);
        atomicAdd(grid + rc.w * gw + rc.z, 1);
        if (!depths) continue;                  // tile-segment sort: no digit histograms
        const uint32_t db = __float_as_uint(__ldg(depths + i));
        atomicAdd(s_hist + 0 * SORT_BINS + (db & 0xffu), touched);
        atomicAdd(s_hist + 1 * SORT_BINS + ((db >> 8) & 0xffu), touched);
        atomicAdd(s_hist + 2 * SORT_BINS + ((db >> 16) & 0xffu), touched);
        atomicAdd(s_hist + 3 * SORT_BINS + (db >> 24), touched);
    }
    __syncthreads();
    if (depths)
        for (int k = threadIdx.x; k < 4 * SORT_BINS; k += COUNT_THREADS)
            if (s_hist[k]) atomicAdd(hist_g + k, s_hist[k]);
    if (in_smem)
        for (int k = threadIdx.x; k < cells; k += COUNT_THREADS)
            if (s_grid[k]) atomicAdd(grid_g + k, s_grid[k]);
}

constexpr int PREP_THREADS = 1024;
constexpr int PREP_VPT = 4;             // values per thread and scan round
constexpr int WORK_BUCKETS = 128;
// bucket 0 = heaviest: 4 buckets per octave of the list length, descending
__device__ __forceinline__ int work_bucket(uint32_t c) {
    if (c == 0) return WORK_BUCKETS - 1;
    const int l2 = 31 - __clz(c);
    const int frac = l2 >= 2 ? (int)((c >> (l2 - 2)) & 3u) : 0;
    return max(0, WORK_BUCKETS - 2 - (l2 * 4 + frac));
}     // difference-array cells integrated in shared memory (covers 1920x1080: 121 x 69)

// block-wide inclusive scan of one value per thread (1024 threads); returns the inclusive value, total via s_ws[31]
__device__ __forceinline__ uint32_t prep_incl_scan(uint32_t c, uint32_t *s_ws) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += n;
    }
    if (lane == 31) s_ws[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = s_ws[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t n = __shfl_up_sync(0xffffffffu, w, d);
            if (lane >= d) w += n;
        }
        s_ws[lane] = w;
    }
    __syncthreads();
    return (warp ? s_ws[warp - 1] : 0) + incl;
}

__global__ void __launch_bounds__(PREP_THREADS) binning_prep_kernel(int gx, int gy, int passes, int end_bit, int nblocks,
                                                                    int32_t *__restrict__ grid_g, uint2 *__restrict__ ranges,
                                                                    uint32_t *__restrict__ hist, uint32_t *__restrict__ block_sums,
                                                                    uint32_t *__restrict__ R_out, uint32_t *__restrict__ tile_order,
                                                                    uint32_t *host_rb, uint32_t host_seq) {
    extern __shared__ int32_t s_grid[];
    __shared__ uint32_t s_h[SORT_MAX_PASSES * SORT_BINS];
    __shared__ uint32_t s_ws[32];
    __shared__ uint32_t s_carry;
    __shared__ uint32_t s_bucket[WORK_BUCKETS];
    __shared__ uint32_t s_longest;
    const int gw = gx + 1, cells = gw * (gy + 1), tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < WORK_BUCKETS) s_bucket[tid] = 0;
    if (tid == 0) s_longest = 0;
    pdl_wait();                                  // launched behind binning_count
    const bool in_smem = cells <= PREP_GRID_SMEM;
    int32_t *grid = in_smem ? s_grid : grid_g;
    if (passes)      // the digit histograms exist only on the onesweep path
        for (int k = tid; k < SORT_MAX_PASSES * SORT_BINS; k += PREP_THREADS) s_h[k] = k < 4 * SORT_BINS ? hist[k] : 0u;
    if (in_smem)
        for (int k = tid; k < cells; k += PREP_THREADS) s_grid[k] = grid_g[k];
    // (a) exclusive scan of the preprocess blocks' instance counts -> per-block key offsets, and R
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < nblocks; base += PREP_THREADS * PREP_VPT) {      // four consecutive values per thread and round
        const int b0 = base + tid * PREP_VPT;
        uint32_t c[PREP_VPT], tsum = 0;
#pragma unroll
        for (int u = 0; u < PREP_VPT; ++u) { c[u] = b0 + u < nblocks ? block_sums[b0 + u] : 0u; tsum += c[u]; }
        uint32_t run = s_carry + prep_incl_scan(tsum, s_ws) - tsum;
#pragma unroll
        for (int u = 0; u < PREP_VPT; ++u) { if (b0 + u < nblocks) block_sums[b0 + u] = run; run += c[u]; }
        __syncthreads();
        if (tid == 0) s_carry += s_ws[31];
        __syncthreads();
    }
    if (tid == 0) { *R_out = s_carry; s_carry = 0; }
    // (b) 2-D inclusive prefix sum of the difference array: along x per row, then along y per column
    for (int y = warp; y <= gy; y += PREP_THREADS / 32) {        // a warp per row, 32 cells per shuffle scan
        int carry = 0;
        for (int x0 = 0; x0 <= gx; x0 += 32) {
            const int x = x0 + lane;
            int v = x <= gx ? grid[y * gw + x] : 0;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int n = __shfl_up_sync(0xffffffffu, v, d);
                if (lane >= d) v += n;
            }
            v += carry;
            if (x <= gx) grid[y * gw + x] = v;
            carry = __shfl_sync(0xffffffffu, v, 31);
        }
    }
    __syncthreads();
    for (int x = tid; x <= gx; x += PREP_THREADS) {
        int run = 0;
        for (int y = 0; y <= gy; ++y) { run += grid[y * gw + x]; grid[y * gw + x] = run; }
    }
    __syncthreads();
    // (c) exclusive scan of the counts in tile order -> ranges; histograms of the key digits that hold the tile id
    const int tiles = gx * gy;
    for (int base = 0; base < tiles; base += PREP_THREADS * PREP_VPT) {
        const int t0 = base + tid * PREP_VPT;
        uint32_t c[PREP_VPT], tsum = 0;
#pragma unroll
        for (int u = 0; u < PREP_VPT; ++u) {
            const int t = t0 + u;
            c[u] = t < tiles ? (uint32_t)grid[(t / gx) * gw + (t % gx)] : 0u;
            tsum += c[u];
        }
        uint32_t start = s_carry + prep_incl_scan(tsum, s_ws) - tsum;
#pragma unroll
        for (int u = 0; u < PREP_VPT; ++u) {
            const int t = t0 + u;
            if (t < tiles) {
                ranges[t] = c[u] ? make_uint2(start, start + c[u]) : make_uint2(0u, 0u);
                atomicAdd(&s_bucket[work_bucket(c[u])], 1u);
                if (c[u] >= 1536u) atomicMax(&s_longest, c[u]);       // only long lists matter to the reader (tile_sort's class choice)
                if (c[u]) {
                    for (int q = 4; q < passes; ++q) {
                        const int shift = 8 * (q - 4), nb = min(8, end_bit - 8 * q);
                        atomicAdd(&s_h[q * SORT_BINS + (((uint32_t)t >> shift) & ((1u << nb) - 1u))], c[u]);
                    }
                }
            }
            start += c[u];
        }
        __syncthreads();
        if (tid == 0) s_carry += s_ws[31];
        __syncthreads();
    }
    // (c') tile launch order for the blend kernels: heaviest lists first (longest-processing-time-first keeps the tail of
    // the launch short); a counting sort over ~quarter-octave buckets of the list length is plenty
    if (warp == 0) {         // exclusive scan of the 128 bucket counts: 4 per lane
        uint32_t n[WORK_BUCKETS / 32], sum = 0;
#pragma unroll
        for (int k = 0; k < WORK_BUCKETS / 32; ++k) { n[k] = s_bucket[lane * (WORK_BUCKETS / 32) + k]; sum += n[k]; }
        uint32_t incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += v;
        }
        uint32_t run = incl - sum;
#pragma unroll
        for (int k = 0; k < WORK_BUCKETS / 32; ++k) { s_bucket[lane * (WORK_BUCKETS / 32) + k] = run; run += n[k]; }
        if (lane == 0) {
            R_out[1] = s_longest;            // longest tile list (0 if below 1024), read back together with R
            R_out[2] = 0u;                   // visible-Gaussian counter of the key emission that follows
            if (host_rb) {
                // the host's copy of (R, longest list), written straight into its pinned memory and published by a sequence
                // number it polls: no copy or event between this kernel and the key emission, which can therefore be a
                // programmatic dependent launch
                volatile uint32_t *h = host_rb;
                h[0] = *R_out; h[1] = s_longest;
                __threadfence_system();
                h[2] = host_seq;
            }
        }
    }
    __syncthreads();
    for (int t = tid; t < tiles; t += PREP_THREADS) {
        const uint32_t c = (uint32_t)grid[(t / gx) * gw + (t % gx)];
        tile_order[atomicAdd(&s_bucket[work_bucket(c)], 1u)] = (uint32_t)t;
    }
    // (d) exclusive scan of every digit histogram: one warp per pass, 8 bins per lane
    if (warp < passes) {     // (after the tile loop's last barrier: s_h is complete)
        uint32_t v[8], sum = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) { v[k] = s_h[warp * SORT_BINS + lane * 8 + k]; sum += v[k]; }
        uint32_t incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t n = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += n;
        }
        uint32_t run = incl - sum;
#pragma unroll
        for (int k = 0; k < 8; ++k) { hist[warp * SORT_BINS + lane * 8 + k] = run; run += v[k]; }
    }
}

// end_bit > 0: global onesweep sort follows (digit histograms are prepared); end_bit == 0: tile-segment sort (cursors)
int launch_binning_prep(int P, int W, int H, int end_bit, const GeomPtrs &g, const ImgPtrs &im, uint32_t *host_rb, uint32_t host_seq, cudaStream_t s) {
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    const int passes = (end_bit + 7) / 8;
    const bool onesweep = end_bit > 0;
    const int cells = (gx + 1) * (gy + 1);
    const size_t dyn = cells <= PREP_GRID_SMEM ? (size_t)cells * sizeof(int32_t) : 0;
    {   // (fusing the four corner updates into preprocess_forward as global REDs was measured: 0.031 -> 0.070 ms there, the
        // ~2 k cells serialise in L2; these few CTAs with shared-memory tables cost 0.010 ms)
        const int blocks = min(148, ceil_div(P, COUNT_THREADS * 2));
        LVDGS_PRE(s);
        LVDGS_CHECK(launch_after_kernel(binning_count_kernel, dim3(blocks), dim3(COUNT_THREADS), 4 * SORT_BINS * sizeof(uint32_t) + dyn, s,
                                        P, gx, gy, g.tiles_touched, g.rect, onesweep ? g.depths : nullptr, im.tile_grid, im.sort_hist));
        LVDGS_LAUNCHED(s, "binning_count");
    }
    LVDGS_PRE(s);
    LVDGS_CHECK(launch_after_kernel(binning_prep_kernel, dim3(1), dim3(PREP_THREADS), dyn, s, gx, gy, passes, end_bit, ceil_div(P, PRE_THREADS),
                                    im.tile_grid, im.ranges, im.sort_hist, g.block_sums, g.num_instances, im.tile_order, host_rb, host_seq));
    LVDGS_LAUNCHED(s, "binning_prep");
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// K3: key emission.  Block b owns Gaussians [256b, 256b+256); its instances are the contiguous span
// [offsets[256b-1], offsets[256b+255]).  Thread t writes instances t, t+256, ... of that span, locating the
// owning Gaussian by binary search over the block's offsets in shared memory.  Emission order inside a
// Gaussian is y-major, x-minor -- the order the reference's per-thread loop produces -- so ties in the
// stable sort resolve identically.
// ---------------------------------------------------------------------------------------------------------
constexpr int EMIT_THREADS = 256;
static_assert(EMIT_THREADS == PRE_THREADS, "emit blocks must own the same Gaussians as preprocess blocks (block_sums)");
static_assert(PRE_THREADS == SORT_BINS, "s_hist3 is indexed by threadIdx");

// BUCKET = true (tile-segment sort, tile_sort.cu): the instance goes to the next free slot of its tile's segment as the
// word (depth bits << 32 | Gaussian); slot order is arbitrary, the per-tile sort makes the result unique.
template <bool BUCKET>
__global__ void __launch_bounds__(EMIT_THREADS) emit_keys_kernel(int P, int gx, uint32_t capacity, const uint32_t *__restrict__ block_offsets,
                                                                 const uint32_t *__restrict__ tiles_touched, uint32_t *__restrict__ offsets,
                                                                 const short4 *__restrict__ rects,
                                                                 const float *__restrict__ depths,
                                                                 uint64_t *__restrict__ keys, uint32_t *__restrict__ vals,
                                                                 uint32_t *__restrict__ tile_cursor, const uint2 *__restrict__ ranges,
                                                                 uint32_t *__restrict__ visible_list, uint32_t *__restrict__ num_visible, int32_t *__restrict__ sel_out) {
    __shared__ uint32_t s_end[EMIT_THREADS];      // inclusive offsets of this block's Gaussians
    __shared__ uint32_t s_vcnt[EMIT_THREADS / 32];
    __shared__ uint32_t s_vbase;
    __shared__ short4 s_rect[EMIT_THREADS];
    __shared__ uint32_t s_depth[EMIT_THREADS];
    __shared__ uint32_t s_wsum[EMIT_THREADS / 32];
    const int g0 = blockIdx.x * EMIT_THREADS;
    const int i = g0 + threadIdx.x;
    pdl_wait();                                  // launched behind binning_prep when the host polls for R (no copy in between)
    if (BUCKET && sel_out && i == 0) *sel_out = 1;              // the tile-segment sort leaves its result in keys[1] / vals[1]
    const uint32_t span_begin = block_offsets[blockIdx.x];      // exclusive prefix over the preceding blocks (binning_prep)
    // K2, second half: inclusive scan of this block's tile counts
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = i < P ? tiles_touched[i] : 0u;
    // the visible Gaussians (at least one tile) are also listed compactly for the preprocess backward: ranks inside the
    // block by ballot, the block's base by one atomic (block order is arbitrary, the backward does not care)
    const uint32_t vis_ballot = __ballot_sync(0xffffffffu, incl != 0u);
    const bool is_vis = incl != 0u;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += n;
    }
    if (lane == 31) s_wsum[warp] = incl;
    if (lane == 0) s_vcnt[warp] = __popc(vis_ballot);
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
#pragma unroll
        for (int k = 0; k < EMIT_THREADS / 32; ++k) { const uint32_t c = s_vcnt[k]; s_vcnt[k] = t; t += c; }
        s_vbase = t ? atomicAdd(num_visible, t) : 0u;
    }
    uint32_t wbase = 0;
#pragma unroll
    for (int k = 0; k < EMIT_THREADS / 32; ++k) wbase += k < warp ? s_wsum[k] : 0u;
    incl += span_begin + wbase;
    if (i < P) {
        offsets[i] = incl;                                      // point_offsets, kept for inspection / parity tests
        s_end[threadIdx.x] = incl;
        s_rect[threadIdx.x] = rects[i];
        s_depth[threadIdx.x] = __float_as_uint(depths[i]);
    } else {
        s_end[threadIdx.x] = 0xffffffffu;
    }
    __syncthreads();
    if (is_vis) visible_list[s_vbase + s_vcnt[warp] + __popc(vis_ballot & ((1u << lane) - 1u))] = (uint32_t)i;
    const int last = min(EMIT_THREADS, P - g0) - 1;
    const uint32_t span_end = BUCKET ? s_end[last] : min(s_end[last], capacity);      // never write past the arena the launch was sized for
    // (four instances per thread and round with their atomics in flight together was measured slower, 0.049 vs 0.045 ms: the
    // kernel is bound by the L2 atomic units' throughput, not by the round trip)
    for (uint32_t r = span_begin + threadIdx.x; r < span_end; r += EMIT_THREADS) {
        // smallest j with s_end[j] > r
        int lo = 0, hi = last;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (s_end[mid] > r) hi = mid; else lo = mid + 1;
        }
        const uint32_t begin = lo ? s_end[lo - 1] : span_begin;
        const short4 rc = s_rect[lo];
        const uint32_t local = r - begin;
        const uint32_t w = (uint32_t)(rc.z - rc.x);
        const uint32_t yy = local / w, xx = local - yy * w;
        const uint32_t tile = (uint32_t)(rc.y + (int)yy) * (uint32_t)gx + (uint32_t)(rc.x + (int)xx);
        if (BUCKET) {
            const uint32_t pos = __ldg(&ranges[tile].x) + atomicAdd(tile_cursor + (size_t)tile * CURSOR_STRIDE, 1u);
            if (pos < capacity) keys[pos] = ((uint64_t)s_depth[lo] << 32) | (uint32_t)(g0 + lo);
        } else {
            keys[r] = ((uint64_t)tile << 32) | s_depth[lo];
            vals[r] = (uint32_t)(g0 + lo);
        }
    }
}

int launch_emit_keys(int P, int W, int H, const GeomPtrs &g, int64_t capacity, uint64_t *keys, uint32_t *vals,
                     uint32_t *tile_cursor, const uint2 *ranges, int32_t *sel_out, cudaStream_t s) {
    (void)H;
    const int gx = (W + TILE - 1) / TILE;
    const uint32_t cap = (uint32_t)min(capacity, (int64_t)0xffffffffll);
    LVDGS_PRE(s);
    if (tile_cursor)
        LVDGS_CHECK(launch_after_kernel(emit_keys_kernel<true>, dim3(ceil_div(P, EMIT_THREADS)), dim3(EMIT_THREADS), 0, s, P, gx, cap, g.block_sums, g.tiles_touched,
                                        g.point_offsets, g.rect, g.depths, keys, vals, tile_cursor, ranges, g.visible_list, g.num_instances + 2, sel_out));
    else
        emit_keys_kernel<false><<<ceil_div(P, EMIT_THREADS), EMIT_THREADS, 0, s>>>(P, gx, cap, g.block_sums, g.tiles_touched, g.point_offsets, g.rect, g.depths, keys, vals, nullptr, nullptr, g.visible_list, g.num_instances + 2, nullptr);
    LVDGS_LAUNCHED(s, "emit_keys");
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// K10: markVisible
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PRE_THREADS) mark_visible_kernel(int P, const float *__restrict__ means3D,
                                                                   const float *__restrict__ view,
                                                                   uint8_t *__restrict__ present) {
    __shared__ float stage[PRE_THREADS * 3];
    const float3 p = load3_staged(means3D, P, stage, blockIdx.x);
    const int i = blockIdx.x * PRE_THREADS + threadIdx.x;
    if (i >= P) return;
    const float tz = __fadd_rn(dot3c(__ldg(view + 2), p.x, __ldg(view + 6), p.y, __ldg(view + 10), p.z), __ldg(view + 14));
    present[i] = tz > 0.2f ? 1 : 0;
}

int launch_mark_visible(int P, const float *means3D, const float *view, uint8_t *present, cudaStream_t s) {
    if (P <= 0) return 0;
    LVDGS_PRE(s);
    mark_visible_kernel<<<ceil_div(P, PRE_THREADS), PRE_THREADS, 0, s>>>(P, means3D, view, present);
    LVDGS_LAUNCHED(s, "mark_visible");
    return 0;
}

}  // namespace lvdgs
