// preprocess_backward.cu -- K8 + K9 + K11 fused: conic -> cov2D -> cov3D/mean gradients, mean2D/depth -> mean
// gradients, SH backward, cov3D -> scale/rotation, and the camera-pose gradient dL/dtau = (rho, theta) of the
// pose-aware fork (consumer: utils/pose_utils.py:70-87 update_pose, tau = [rho; theta],
// T_new = Exp(tau) T_w2c).  One thread per Gaussian, one kernel instead of the reference's two, and the
// per-Gaussian pose gradients are block-reduced and added into dL_dtau_sum[6] here, replacing the
// `grad_tau.view(-1,6).sum(0)` torch reduction (K11) and, when the caller does not ask for it, the [P,6] array.
#include "common.cuh"

namespace lvdgs {

#ifndef LVDGS_PB_THREADS
#define LVDGS_PB_THREADS 128
#endif
#ifndef LVDGS_PB_MINBLOCKS
#define LVDGS_PB_MINBLOCKS 8      // <= 64 registers: the kernel is latency-bound, occupancy measured to pay (56 -> 43 us)
#endif
constexpr int PB_THREADS = LVDGS_PB_THREADS;

__device__ __constant__ float B_SH_C0 = 0.28209479177387814f;
__device__ __constant__ float B_SH_C1 = 0.4886025119029199f;
__device__ __constant__ float B_SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                            -1.0925484305920792f, 0.5462742152960396f};
__device__ __constant__ float B_SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                            0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                                            -0.5900435899266435f};

struct PbArgs {
    int P, D, M, W, H, flags;
    float tanfovx, tanfovy, fx, fy, mod;
    const float *means3D, *shs, *scales, *rotations, *cov3D_precomp, *view, *proj, *proj_raw, *campos;
    const int32_t *radii;
    const uint8_t *clamped;
    const float4 *conic_opacity;
    const float *acc;          // accumulator rows of the blend backward
    bool zero_culled_means2D;  // compact mode: clear dL_dmeans2D rows of culled Gaussians here (no memset launch)
    float *dL_dmeans2D, *dL_dcolors, *dL_dopacity, *dL_dmeans3D, *dL_dcov3D, *dL_dsh, *dL_dscales, *dL_drots,
        *dL_dtau, *dL_dtau_sum;
    // compact mode (accumulate / pose-only / pre-zeroed outputs, no per-Gaussian tau): thread j handles Gaussian visible_list[j], j < *num_visible.
    // Only ~40-60 % of a map is in view, so the kernel runs that fraction of the warps, all lanes live; culled rows are
    // never touched (accumulate mode leaves them as they are; their dL_dmeans2D rows are cleared by the block that owns
    // their index range)
    const uint32_t *visible_list, *num_visible;
};

__device__ __forceinline__ float3 cross3(float3 a, float3 b) {
    return make_float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

__device__ __forceinline__ void quat_R(float4 q, float R[9]) {
    const float r = q.x, x = q.y, y = q.z, z = q.w;
    R[0] = 1.f - 2.f * (y * y + z * z); R[1] = 2.f * (x * y - r * z); R[2] = 2.f * (x * z + r * y);
    R[3] = 2.f * (x * y + r * z); R[4] = 1.f - 2.f * (x * x + z * z); R[5] = 2.f * (y * z - r * x);
    R[6] = 2.f * (x * z - r * y); R[7] = 2.f * (y * z + r * x); R[8] = 1.f - 2.f * (x * x + y * y);
}

__global__ void __launch_bounds__(PB_THREADS, LVDGS_PB_MINBLOCKS) preprocess_backward_kernel(const PbArgs a) {
    __shared__ CameraConst cam;
    __shared__ float s_praw[16];
    __shared__ float s_tau[PB_THREADS / 32][6];
    if (threadIdx.x < 16) cam.view[threadIdx.x] = __ldg(a.view + threadIdx.x);
    else if (threadIdx.x < 32) cam.proj[threadIdx.x - 16] = __ldg(a.proj + threadIdx.x - 16);
    else if (threadIdx.x < 35) cam.campos[threadIdx.x - 32] = __ldg(a.campos + threadIdx.x - 32);
    else if (threadIdx.x >= 64 && threadIdx.x < 80) s_praw[threadIdx.x - 64] = __ldg(a.proj_raw + threadIdx.x - 64);
    int i = blockIdx.x * PB_THREADS + threadIdx.x;
    pdl_wait();                                  // launched behind the blend backward (programmatic dependent launch)
    if (a.visible_list) {
        // the per-view screen-space gradient of a culled Gaussian is zero: every block clears the culled rows of its own
        // index range (the listed, i.e. visible, rows are written by whichever block walks them -- disjoint sets)
        if (a.zero_culled_means2D && i < a.P && a.radii[i] <= 0) {
            float *row = a.dL_dmeans2D + 3 * (size_t)i;
            row[0] = 0.f; row[1] = 0.f; row[2] = 0.f;
        }
        const uint32_t nv = __ldg(a.num_visible);
        if ((uint32_t)(blockIdx.x * PB_THREADS) >= nv) return;               // block-uniform
        i = (uint32_t)i < nv ? (int)__ldg(a.visible_list + i) : a.P;          // a.P: out of range, the lane idles
    }
    __syncthreads();
    const float *V = cam.view, *Pj = cam.proj;

    float dmean[3] = {0.f, 0.f, 0.f}, dcov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, tau[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float dscale[3] = {0.f, 0.f, 0.f}, drot[4] = {0.f, 0.f, 0.f, 0.f};
    float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0, r2 = r0;
    const bool live = i < a.P && (a.visible_list != nullptr || a.radii[i] > 0);     // a listed Gaussian is visible
    // tracking (utils/slam_frontend.py:1468-1521) optimises the camera only: no Gaussian parameter needs a gradient
    const bool pose_only = (a.flags & LVDGS_FLAG_POSE_ONLY) != 0;
    if (live) {
        // moments of the blend backward -> dL_dmean2D (NDC units), dL_dconic, dL_dopacity (see ACC_STRIDE, common.cuh)
        const float4 *row = reinterpret_cast<const float4 *>(a.acc + (size_t)i * ACC_STRIDE);
        const float4 m0 = row[0], m1 = row[1];
        r2 = row[2];

        const float4 co = a.conic_opacity[i];
        const float o = co.w;
        r0.x = -0.5f * (float)a.W * o * (co.x * m0.x + co.y * m0.y);
        r0.y = -0.5f * (float)a.H * o * (co.y * m0.x + co.z * m0.y);
        r0.z = -0.5f * o * m0.z;       // dL_dconic xx
        r0.w = -0.5f * o * m0.w;       // dL_dconic xy
        r1.x = -0.5f * o * m1.x;       // dL_dconic yy
        r1.y = m1.y;                   // dL_dopacity
        r1.z = m1.z;                   // dL_ddepth
    }
    if (live) {
        const float x = a.means3D[3 * (size_t)i], y = a.means3D[3 * (size_t)i + 1], z = a.means3D[3 * (size_t)i + 2];
        const float3 pc = make_float3(V[0] * x + V[4] * y + V[8] * z + V[12], V[1] * x + V[5] * y + V[9] * z + V[13],
                                      V[2] * x + V[6] * y + V[10] * z + V[14]);
        // cov3D (recomputed; the forward does not store it)
        float c3[6];
        float Rq[9];
        float4 q = make_float4(1.f, 0.f, 0.f, 0.f);
        float s[3] = {0.f, 0.f, 0.f};
        if (a.cov3D_precomp) {
#pragma unroll
            for (int k = 0; k < 6; ++k) c3[k] = a.cov3D_precomp[(size_t)i * 6 + k];
        } else {
            q = __ldg(reinterpret_cast<const float4 *>(a.rotations) + i);
            quat_R(q, Rq);
#pragma unroll
            for (int k = 0; k < 3; ++k) s[k] = a.mod * a.scales[3 * (size_t)i + k];
            float A[9];
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int k = 0; k < 3; ++k) A[r * 3 + k] = Rq[r * 3 + k] * s[k];
            c3[0] = A[0] * A[0] + A[1] * A[1] + A[2] * A[2];
            c3[1] = A[0] * A[3] + A[1] * A[4] + A[2] * A[5];
            c3[2] = A[0] * A[6] + A[1] * A[7] + A[2] * A[8];
            c3[3] = A[3] * A[3] + A[4] * A[4] + A[5] * A[5];
            c3[4] = A[3] * A[6] + A[4] * A[7] + A[5] * A[8];
            c3[5] = A[6] * A[6] + A[7] * A[7] + A[8] * A[8];
        }
        // ---- A.5.1 : EWA recompute ----
        const float limx = 1.3f * a.tanfovx, limy = 1.3f * a.tanfovy;
        const float txtz = pc.x / pc.z, tytz = pc.y / pc.z;
        const float xmul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
        const float ymul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
        const float3 t = make_float3(fminf(limx, fmaxf(-limx, txtz)) * pc.z, fminf(limy, fmaxf(-limy, tytz)) * pc.z, pc.z);
        const float tzi = 1.f / t.z, tz2 = tzi * tzi, tz3 = tz2 * tzi;
        const float J00 = a.fx * tzi, J02 = -a.fx * t.x * tz2, J11 = a.fy * tzi, J12 = -a.fy * t.y * tz2;
        float R0[3], R1[3], R2[3], m0[3], m1[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            R0[k] = V[4 * k]; R1[k] = V[4 * k + 1]; R2[k] = V[4 * k + 2];
            m0[k] = J00 * R0[k] + J02 * R2[k];
            m1[k] = J11 * R1[k] + J12 * R2[k];
        }
        float u0[3], u1[3];
        u0[0] = c3[0] * m0[0] + c3[1] * m0[1] + c3[2] * m0[2];
        u0[1] = c3[1] * m0[0] + c3[3] * m0[1] + c3[4] * m0[2];
        u0[2] = c3[2] * m0[0] + c3[4] * m0[1] + c3[5] * m0[2];
        u1[0] = c3[0] * m1[0] + c3[1] * m1[1] + c3[2] * m1[2];
        u1[1] = c3[1] * m1[0] + c3[3] * m1[1] + c3[4] * m1[2];
        u1[2] = c3[2] * m1[0] + c3[4] * m1[1] + c3[5] * m1[2];
        const float ca = m0[0] * u0[0] + m0[1] * u0[1] + m0[2] * u0[2] + 0.3f;
        const float cb = m0[0] * u1[0] + m0[1] * u1[1] + m0[2] * u1[2];
        const float cc = m1[0] * u1[0] + m1[1] * u1[1] + m1[2] * u1[2] + 0.3f;
        const float gxc = r0.z, gyc = r0.w, gzc = r1.x;      // dL_dconic xx, xy, yy
        const float denom = ca * cc - cb * cb;
        const float k2 = 1.f / (denom * denom + 0.0000001f);
        float da = 0.f, db = 0.f, dc = 0.f;
        if (k2 != 0.f) {
            da = k2 * (-cc * cc * gxc + 2.f * cb * cc * gyc + (denom - ca * cc) * gzc);
            dc = k2 * (-ca * ca * gzc + 2.f * ca * cb * gyc + (denom - ca * cc) * gxc);
            db = k2 * 2.f * (cb * cc * gxc - (denom + 2.f * cb * cb) * gyc + ca * cb * gzc);
            dcov[0] = m0[0] * m0[0] * da + m0[0] * m1[0] * db + m1[0] * m1[0] * dc;
            dcov[3] = m0[1] * m0[1] * da + m0[1] * m1[1] * db + m1[1] * m1[1] * dc;
            dcov[5] = m0[2] * m0[2] * da + m0[2] * m1[2] * db + m1[2] * m1[2] * dc;
            dcov[1] = 2.f * m0[0] * m0[1] * da + (m0[0] * m1[1] + m0[1] * m1[0]) * db + 2.f * m1[0] * m1[1] * dc;
            dcov[2] = 2.f * m0[0] * m0[2] * da + (m0[0] * m1[2] + m0[2] * m1[0]) * db + 2.f * m1[0] * m1[2] * dc;
            dcov[4] = 2.f * m0[2] * m0[1] * da + (m0[1] * m1[2] + m0[2] * m1[1]) * db + 2.f * m1[1] * m1[2] * dc;
        }
        float dm0[3], dm1[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            dm0[k] = 2.f * u0[k] * da + u1[k] * db;
            dm1[k] = 2.f * u1[k] * dc + u0[k] * db;
        }
        const float dJ00 = R0[0] * dm0[0] + R0[1] * dm0[1] + R0[2] * dm0[2];
        const float dJ02 = R2[0] * dm0[0] + R2[1] * dm0[1] + R2[2] * dm0[2];
        const float dJ11 = R1[0] * dm1[0] + R1[1] * dm1[1] + R1[2] * dm1[2];
        const float dJ12 = R2[0] * dm1[0] + R2[1] * dm1[1] + R2[2] * dm1[2];
        float3 dt;
        dt.x = xmul * -a.fx * tz2 * dJ02;
        dt.y = ymul * -a.fy * tz2 * dJ12;
        dt.z = -a.fx * tz2 * dJ00 - a.fy * tz2 * dJ11 + (2.f * a.fx * t.x) * tz3 * dJ02 + (2.f * a.fy * t.y) * tz3 * dJ12;
#pragma unroll
        for (int k = 0; k < 3; ++k) dmean[k] = R0[k] * dt.x + R1[k] * dt.y + R2[k] * dt.z;
        {   // pose through t (clamped t, as in the recomputed forward) and through the view rotation's columns
            const float3 cr = cross3(t, dt);
            tau[0] += dt.x; tau[1] += dt.y; tau[2] += dt.z;
            tau[3] += cr.x; tau[4] += cr.y; tau[5] += cr.z;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float3 col = make_float3(R0[k], R1[k], R2[k]);
                const float3 g = make_float3(J00 * dm0[k], J11 * dm1[k], J02 * dm0[k] + J12 * dm1[k]);
                const float3 c = cross3(col, g);
                tau[3] += c.x; tau[4] += c.y; tau[5] += c.z;
            }
        }
        // ---- A.5.2 : mean2D and depth ----
        const float hx = Pj[0] * x + Pj[4] * y + Pj[8] * z + Pj[12];
        const float hy = Pj[1] * x + Pj[5] * y + Pj[9] * z + Pj[13];
        const float hw = Pj[3] * x + Pj[7] * y + Pj[11] * z + Pj[15];
        const float mw = 1.f / (hw + 0.0000001f);
        const float mul1 = hx * mw * mw, mul2 = hy * mw * mw;
        const float g2x = r0.x, g2y = r0.y;
        dmean[0] += (Pj[0] * mw - Pj[3] * mul1) * g2x + (Pj[1] * mw - Pj[3] * mul2) * g2y;
        dmean[1] += (Pj[4] * mw - Pj[7] * mul1) * g2x + (Pj[5] * mw - Pj[7] * mul2) * g2y;
        dmean[2] += (Pj[8] * mw - Pj[11] * mul1) * g2x + (Pj[9] * mw - Pj[11] * mul2) * g2y;
        {
            const float al = mw, be = -hx * mw * mw, ga = -hy * mw * mw;
            const float pa = s_praw[0], pb = s_praw[5], pe = s_praw[11];
            float3 d1 = make_float3(al * pa, 0.f, be * pe), d2 = make_float3(0.f, al * pb, ga * pe);
            if (a.flags & LVDGS_FLAG_EXACT_PP) { d1.z += al * s_praw[8]; d2.z += al * s_praw[9]; }
            const float3 v = make_float3(g2x * d1.x + g2y * d2.x, g2x * d1.y + g2y * d2.y, g2x * d1.z + g2y * d2.z);
            const float3 cr = cross3(pc, v);
            tau[0] += v.x; tau[1] += v.y; tau[2] += v.z;
            tau[3] += cr.x; tau[4] += cr.y; tau[5] += cr.z;
        }
        {
            const float dz = r1.z;                       // dL_ddepth
            dmean[0] += dz * V[2]; dmean[1] += dz * V[6]; dmean[2] += dz * V[10];
            tau[2] += dz;
            tau[3] += dz * pc.y;
            tau[4] += dz * -pc.x;
        }
        // ---- SH backward ----
        if (!pose_only && a.dL_dsh && a.shs) {
            const float *sh = a.shs + (size_t)i * a.M * 3;
            float *dsh = a.dL_dsh + (size_t)i * a.M * 3;
            const uint8_t cl = a.clamped[i];
            const float dRGB[3] = {(cl & 1) ? 0.f : r2.x, (cl & 2) ? 0.f : r2.y, (cl & 4) ? 0.f : r2.z};
            const int ncoef = (a.D + 1) * (a.D + 1);
            const bool accum_sh = (a.flags & LVDGS_FLAG_ACCUMULATE) != 0;
            if (!accum_sh) for (int k = ncoef * 3; k < a.M * 3; ++k) dsh[k] = 0.f;
            if (a.D == 0) {
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) { if (accum_sh) dsh[ch] += B_SH_C0 * dRGB[ch]; else dsh[ch] = B_SH_C0 * dRGB[ch]; }
            } else {
                const float3 dorig = make_float3(x - cam.campos[0], y - cam.campos[1], z - cam.campos[2]);
                const float s2 = dorig.x * dorig.x + dorig.y * dorig.y + dorig.z * dorig.z;
                const float inv = rsqrtf(s2);
                const float sx = dorig.x * inv, sy = dorig.y * inv, sz = dorig.z * inv;
                float ddir[3] = {0.f, 0.f, 0.f};
                for (int ch = 0; ch < 3; ++ch) {
#define SHC(k) sh[(k) * 3 + ch]
#define DSH(k) dshv[k]
                    float dshv[16];
                    const float g = dRGB[ch];
                    float ddx, ddy, ddz;
                    DSH(0) = B_SH_C0 * g;
                    DSH(1) = -B_SH_C1 * sy * g; DSH(2) = B_SH_C1 * sz * g; DSH(3) = -B_SH_C1 * sx * g;
                    ddx = -B_SH_C1 * SHC(3); ddy = -B_SH_C1 * SHC(1); ddz = B_SH_C1 * SHC(2);
                    if (a.D > 1) {
                        const float xx = sx * sx, yy = sy * sy, zz = sz * sz, xy = sx * sy, yz = sy * sz, xz = sx * sz;
                        DSH(4) = B_SH_C2[0] * xy * g; DSH(5) = B_SH_C2[1] * yz * g;
                        DSH(6) = B_SH_C2[2] * (2.f * zz - xx - yy) * g; DSH(7) = B_SH_C2[3] * xz * g;
                        DSH(8) = B_SH_C2[4] * (xx - yy) * g;
                        ddx += B_SH_C2[0] * sy * SHC(4) + B_SH_C2[2] * 2.f * -sx * SHC(6) + B_SH_C2[3] * sz * SHC(7) + B_SH_C2[4] * 2.f * sx * SHC(8);
                        ddy += B_SH_C2[0] * sx * SHC(4) + B_SH_C2[1] * sz * SHC(5) + B_SH_C2[2] * 2.f * -sy * SHC(6) + B_SH_C2[4] * 2.f * -sy * SHC(8);
                        ddz += B_SH_C2[1] * sy * SHC(5) + B_SH_C2[2] * 2.f * 2.f * sz * SHC(6) + B_SH_C2[3] * sx * SHC(7);
                        if (a.D > 2) {
                            DSH(9) = B_SH_C3[0] * sy * (3.f * xx - yy) * g;
                            DSH(10) = B_SH_C3[1] * xy * sz * g;
                            DSH(11) = B_SH_C3[2] * sy * (4.f * zz - xx - yy) * g;
                            DSH(12) = B_SH_C3[3] * sz * (2.f * zz - 3.f * xx - 3.f * yy) * g;
                            DSH(13) = B_SH_C3[4] * sx * (4.f * zz - xx - yy) * g;
                            DSH(14) = B_SH_C3[5] * sz * (xx - yy) * g;
                            DSH(15) = B_SH_C3[6] * sx * (xx - 3.f * yy) * g;
                            ddx += B_SH_C3[0] * SHC(9) * 3.f * 2.f * xy + B_SH_C3[1] * SHC(10) * yz + B_SH_C3[2] * SHC(11) * -2.f * xy +
                                   B_SH_C3[3] * SHC(12) * -3.f * 2.f * xz + B_SH_C3[4] * SHC(13) * (-3.f * xx + 4.f * zz - yy) +
                                   B_SH_C3[5] * SHC(14) * 2.f * xz + B_SH_C3[6] * SHC(15) * 3.f * (xx - yy);
                            ddy += B_SH_C3[0] * SHC(9) * 3.f * (xx - yy) + B_SH_C3[1] * SHC(10) * xz +
                                   B_SH_C3[2] * SHC(11) * (-3.f * yy + 4.f * zz - xx) + B_SH_C3[3] * SHC(12) * -3.f * 2.f * yz +
                                   B_SH_C3[4] * SHC(13) * -2.f * xy + B_SH_C3[5] * SHC(14) * -2.f * yz + B_SH_C3[6] * SHC(15) * -3.f * 2.f * xy;
                            ddz += B_SH_C3[1] * SHC(10) * xy + B_SH_C3[2] * SHC(11) * 4.f * 2.f * yz +
                                   B_SH_C3[3] * SHC(12) * 3.f * (2.f * zz - xx - yy) + B_SH_C3[4] * SHC(13) * 4.f * 2.f * xz +
                                   B_SH_C3[5] * SHC(14) * (xx - yy);
                        }
                    }
                    ddir[0] += ddx * g; ddir[1] += ddy * g; ddir[2] += ddz * g;
                    for (int k = 0; k < ncoef; ++k) { if (accum_sh) dsh[k * 3 + ch] += dshv[k]; else dsh[k * 3 + ch] = dshv[k]; }
#undef SHC
#undef DSH
                }
                const float inv32 = inv * inv * inv;
                float dm[3];
                dm[0] = ((s2 - dorig.x * dorig.x) * ddir[0] - dorig.y * dorig.x * ddir[1] - dorig.z * dorig.x * ddir[2]) * inv32;
                dm[1] = (-dorig.x * dorig.y * ddir[0] + (s2 - dorig.y * dorig.y) * ddir[1] - dorig.z * dorig.y * ddir[2]) * inv32;
                dm[2] = (-dorig.x * dorig.z * ddir[0] - dorig.y * dorig.z * ddir[1] + (s2 - dorig.z * dorig.z) * ddir[2]) * inv32;
#pragma unroll
                for (int k = 0; k < 3; ++k) { dmean[k] += dm[k]; tau[k] -= dm[k]; }
            }
        }
        // ---- cov3D -> scale / rotation ----
        if (!pose_only && !a.cov3D_precomp) {
            const float Gm[9] = {dcov[0], 0.5f * dcov[1], 0.5f * dcov[2], 0.5f * dcov[1], dcov[3], 0.5f * dcov[4],
                                 0.5f * dcov[2], 0.5f * dcov[4], dcov[5]};
            float Q[9];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                float ds = 0.f;
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const float dA = 2.f * (Gm[r * 3 + 0] * Rq[0 * 3 + k] + Gm[r * 3 + 1] * Rq[1 * 3 + k] + Gm[r * 3 + 2] * Rq[2 * 3 + k]) * s[k];
                    ds += dA * Rq[r * 3 + k];
                    Q[r * 3 + k] = dA * s[k];
                }
                dscale[k] = ds;
            }
            const float r = q.x, qx = q.y, qy = q.z, qz = q.w;
            drot[0] = 2.f * (-qz * Q[1] + qy * Q[2] + qz * Q[3] - qx * Q[5] - qy * Q[6] + qx * Q[7]);
            drot[1] = 2.f * (qy * Q[1] + qz * Q[2] + qy * Q[3] - 2.f * qx * Q[4] - r * Q[5] + qz * Q[6] + r * Q[7] - 2.f * qx * Q[8]);
            drot[2] = 2.f * (-2.f * qy * Q[0] + qx * Q[1] + r * Q[2] + qx * Q[3] + qz * Q[5] - r * Q[6] + qz * Q[7] - 2.f * qy * Q[8]);
            drot[3] = 2.f * (-2.f * qz * Q[0] - r * Q[1] + qx * Q[2] + r * Q[3] - 2.f * qz * Q[4] + qy * Q[5] + qx * Q[6] + qy * Q[7]);
        }
    } else if (i < a.P && a.dL_dsh && !pose_only && !(a.flags & LVDGS_FLAG_ACCUMULATE)) {
        float *dsh = a.dL_dsh + (size_t)i * a.M * 3;
        for (int k = 0; k < a.M * 3; ++k) dsh[k] = 0.f;
    }
    if (i < a.P) {
        // parameter gradients are either stored or, with LVDGS_FLAG_ACCUMULATE, added to what the caller's buffers
        // hold (the mapping loss is a SUM over keyframe views, utils/slam_backend.py:266,300); per-view outputs
        // (dL_dmeans2D, dL_dtau) are always stored.
        const bool accum = (a.flags & LVDGS_FLAG_ACCUMULATE) != 0;
#define PUT(ptr, idx, v) do { if (accum) (ptr)[idx] += (v); else (ptr)[idx] = (v); } while (0)
        const size_t i3 = 3 * (size_t)i;
        if (a.dL_dmeans2D) { a.dL_dmeans2D[i3] = r0.x; a.dL_dmeans2D[i3 + 1] = r0.y; a.dL_dmeans2D[i3 + 2] = 0.f; }
      if (!pose_only && (live || !accum)) {       // a culled Gaussian adds nothing: in accumulate mode its rows are not touched at all
        if (a.dL_dcolors) { PUT(a.dL_dcolors, i3, r2.x); PUT(a.dL_dcolors, i3 + 1, r2.y); PUT(a.dL_dcolors, i3 + 2, r2.z); }
        PUT(a.dL_dopacity, i, r1.y);
        PUT(a.dL_dmeans3D, i3, dmean[0]); PUT(a.dL_dmeans3D, i3 + 1, dmean[1]); PUT(a.dL_dmeans3D, i3 + 2, dmean[2]);
        if (a.dL_dcov3D) {
#pragma unroll
            for (int k = 0; k < 6; ++k) PUT(a.dL_dcov3D, 6 * (size_t)i + k, dcov[k]);
        }
        if (a.dL_dscales) { PUT(a.dL_dscales, i3, dscale[0]); PUT(a.dL_dscales, i3 + 1, dscale[1]); PUT(a.dL_dscales, i3 + 2, dscale[2]); }
        if (a.dL_drots) {
            float4 *dst = reinterpret_cast<float4 *>(a.dL_drots) + i;
            float4 v = make_float4(drot[0], drot[1], drot[2], drot[3]);
            if (accum) { const float4 o = *dst; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
            *dst = v;
        }
      }
        if (a.dL_dtau) {
#pragma unroll
            for (int k = 0; k < 6; ++k) a.dL_dtau[6 * (size_t)i + k] = tau[k];
        }
    }
    if (a.dL_dtau_sum) {   // block reduction of the pose gradient, then 6 atomics per block
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            float v = tau[k];
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
            if (lane == 0) s_tau[w][k] = v;
        }
        __syncthreads();
        if (threadIdx.x < 6) {
            float v = 0.f;
#pragma unroll
            for (int ww = 0; ww < PB_THREADS / 32; ++ww) v += s_tau[ww][threadIdx.x];
            if (v != 0.f) atomicAdd(a.dL_dtau_sum + threadIdx.x, v);
        }
    }
}

int launch_preprocess_backward(const lvdgs_raster_params &p, const float *means3D, const int32_t *radii,
                               const float *shs, const float *scales, const float *rotations,
                               const float *cov3D_precomp, const float *view, const float *proj,
                               const float *proj_raw, const float *campos, const GeomPtrs &g,
                               const BlendGradPtrs &bgp, bool colors_are_precomp, float *dL_dmeans2D,
                               float *dL_dcolors, float *dL_dopacity, float *dL_dmeans3D, float *dL_dcov3D,
                               float *dL_dsh, float *dL_dscales, float *dL_drots, float *dL_dtau,
                               float *dL_dtau_sum, bool tau_sum_zeroed, cudaStream_t s) {
    PbArgs a;
    a.P = p.P; a.D = p.sh_degree; a.M = p.sh_coeffs; a.W = p.width; a.H = p.height; a.flags = p.flags;
    a.tanfovx = p.tan_fovx; a.tanfovy = p.tan_fovy;
    a.fx = (float)p.width / (2.f * p.tan_fovx); a.fy = (float)p.height / (2.f * p.tan_fovy);
    a.mod = p.scale_modifier;
    a.means3D = means3D; a.shs = colors_are_precomp ? nullptr : shs; a.scales = scales; a.rotations = rotations;
    a.cov3D_precomp = cov3D_precomp; a.view = view; a.proj = proj; a.proj_raw = proj_raw; a.campos = campos;
    a.radii = radii; a.clamped = g.clamped; a.conic_opacity = g.conic_opacity; a.acc = bgp.acc;
    a.dL_dmeans2D = dL_dmeans2D; a.dL_dcolors = dL_dcolors; a.dL_dopacity = dL_dopacity; a.dL_dmeans3D = dL_dmeans3D;
    a.dL_dcov3D = dL_dcov3D; a.dL_dsh = colors_are_precomp ? nullptr : dL_dsh; a.dL_dscales = dL_dscales;
    a.dL_drots = dL_drots; a.dL_dtau = dL_dtau; a.dL_dtau_sum = dL_dtau_sum;
    if (dL_dtau_sum && !tau_sum_zeroed) LVDGS_CHECK(cudaMemsetAsync(dL_dtau_sum, 0, 6 * sizeof(float), s));
    // the culled rows matter only when something dense is written for them: parameter gradients in store mode, dL_dtau
    const bool zeroed = (p.flags & LVDGS_FLAG_ZEROED_OUTPUTS) != 0;
    const bool compact = ((p.flags & LVDGS_FLAG_ACCUMULATE) || (p.flags & LVDGS_FLAG_POSE_ONLY) || zeroed) && !dL_dtau;
    a.visible_list = compact ? g.visible_list : nullptr;
    a.num_visible = g.num_instances + 2;
    a.zero_culled_means2D = compact && dL_dmeans2D && !zeroed;

    LVDGS_PRE(s);
    LVDGS_CHECK(launch_after_kernel(preprocess_backward_kernel, dim3(ceil_div(p.P, PB_THREADS)), dim3(PB_THREADS), 0, s, a));
    LVDGS_LAUNCHED(s, "preprocess_backward");
    return 0;
}

}  // namespace lvdgs
