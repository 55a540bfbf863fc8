/*
 * oracle/raster_oracle.c -- TEST INFRASTRUCTURE ONLY (never shipped, never on the product path).
 *
 * CPU restatement, in plain C, of the tile-based differentiable Gaussian-splatting rasterizer
 * that zwk0901/LVD_GS-SLAM calls through `diff_gaussian_rasterization` (the MonoGS "-w-pose"
 * fork: extra depth / opacity / n_touched outputs and camera-pose gradients).
 *
 * PARITY PARTLY PINNED.  The CUDA source of that plugin is not in /root/reference (it shipped in
 * submodules.zip, listed in /root/reference/.MISSING_LARGE_BLOBS:1) and the reference has no
 * tests or golden vectors for it: the rasterizer's INTERNAL arithmetic
 * (EWA, 0.3 dilation, 3-sigma radius, 1/255 and 1e-4 cut-offs, key layout) is UNPINNED -- restated from the
 * published algorithm.  What IS pinned by reference-held code (round 2): the pose
 * gradient -- this file's analytic dL/dtau equals finite differences taken THROUGH the reference's own
 * SE3_exp (utils/pose_utils.py:56-68), and the chain Camera -> render -> get_loss_tracking -> backward ->
 * Adam -> update_pose with this oracle as the rasterizer reproduces the trajectory recorded by running
 * the reference's Python (tests/golden/reference_pin.npz, tests/test_reference_pin.py).  oracle/build_ref.py
 * compiles the reference's own CUDA into oracle/_ref/ if its sources ever appear.  Anchors on the in-tree call sites:
 *   - argument tuple / output dict keys : utils/slam_backend.py:98-117,184-194 ; utils/slam_frontend.py:1493-1500
 *   - matrix layout (transposed, i.e. column-major flat arrays) : utils/camera_utils.py:106-120
 *   - pose convention  T_new = Exp(tau) * T_w2c , tau = [rho ; theta]  : utils/pose_utils.py:56-87
 *   - depth / opacity are [1,H,W], n_touched is per Gaussian  : utils/slam_utils.py:53-62,107-121 ; utils/slam_backend.py:147,315
 * The analytic backward below is checked against float64 autograd of the forward in
 * tests/test_oracle_autograd.py wherever upstream's backward is the true derivative; the places
 * where upstream deliberately is not are controlled by `flags`.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this file.  Nothing under lvd_gs-slam_b200/ may.
 *
 * Arithmetic: float32 everywhere the CUDA path is float32.  Every multiply-add on the chain that
 * feeds an INTEGER output (radii, tile rect, depth key bits) is written as an explicit fmaf() in a
 * fixed order ("canonical arithmetic") and this file must be compiled with
 * -ffp-contract=off, so that the CUDA kernels -- which spell the same chain with __fmaf_rn /
 * __fmul_rn / __fadd_rn -- can be compared bit for bit.
 *
 * Build: see oracle/Makefile  (gcc -O2 -fopenmp -ffp-contract=off -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define TILE 16
#define FLAG_EXACT_PP 1      /* include Pr[8],Pr[9] in the pose Jacobian (true derivative); upstream omits them (A.6 item 3) */
#define FLAG_OPACITY_GRAD 2  /* propagate dL/d(out_opacity); upstream drops it (A.6 item 1) */

static const float SH_C0 = 0.28209479177387814f;
static const float SH_C1 = 0.4886025119029199f;
static const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                               -1.0925484305920792f, 0.5462742152960396f};
static const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                               0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                               -0.5900435899266435f};

static inline float dot3f(float a0, float b0, float a1, float b1, float a2, float b2) {
    return fmaf(a2, b2, fmaf(a1, b1, a0 * b0));
}

/* Rotation matrix Rq (row-major math matrix) of quaternion q = (r,x,y,z), NOT re-normalised (A.1 step 3). */
static inline void quat_to_R(const float *q, float R[9]) {
    float r = q[0], x = q[1], y = q[2], z = q[3];
    R[0] = fmaf(-2.f, fmaf(z, z, y * y), 1.f);
    R[1] = 2.f * fmaf(x, y, -(r * z));
    R[2] = 2.f * fmaf(x, z, r * y);
    R[3] = 2.f * fmaf(x, y, r * z);
    R[4] = fmaf(-2.f, fmaf(z, z, x * x), 1.f);
    R[5] = 2.f * fmaf(y, z, -(r * x));
    R[6] = 2.f * fmaf(x, z, -(r * y));
    R[7] = 2.f * fmaf(y, z, r * x);
    R[8] = fmaf(-2.f, fmaf(y, y, x * x), 1.f);
}

/* Sigma = Rq diag(s^2) Rq^T, upper triangle (00,01,02,11,12,22).  A.1 step 3. */
static inline void cov3d_from_scale_rot(const float *scale, float mod, const float *q, float c[6]) {
    float R[9];
    quat_to_R(q, R);
    float s0 = mod * scale[0], s1 = mod * scale[1], s2 = mod * scale[2];
    /* A[a][k] = Rq[a][k] * s_k */
    float A[9];
    for (int a = 0; a < 3; ++a) {
        A[a * 3 + 0] = R[a * 3 + 0] * s0;
        A[a * 3 + 1] = R[a * 3 + 1] * s1;
        A[a * 3 + 2] = R[a * 3 + 2] * s2;
    }
    c[0] = dot3f(A[0], A[0], A[1], A[1], A[2], A[2]);
    c[1] = dot3f(A[0], A[3], A[1], A[4], A[2], A[5]);
    c[2] = dot3f(A[0], A[6], A[1], A[7], A[2], A[8]);
    c[3] = dot3f(A[3], A[3], A[4], A[4], A[5], A[5]);
    c[4] = dot3f(A[3], A[6], A[4], A[7], A[5], A[8]);
    c[5] = dot3f(A[6], A[6], A[7], A[7], A[8], A[8]);
}

typedef struct {
    float tx, ty, tz;     /* clamped camera-space point used by the EWA Jacobian */
    float xmul, ymul;     /* 1 if not clamped else 0 */
    float J00, J02, J11, J12;
    float m0[3], m1[3];   /* rows of J*R */
    float u0[3], u1[3];   /* Sigma*m0, Sigma*m1 */
    float a, b, c;        /* dilated 2D covariance */
} Ewa;

/* A.1 step 4.  view = flat column-major 4x4 (torch row-major memory of W2C^T). */
static inline void ewa_project(float tx, float ty, float tz, float fx, float fy, float tanfovx,
                               float tanfovy, const float *cov3D, const float *view, Ewa *e) {
    float limx = 1.3f * tanfovx, limy = 1.3f * tanfovy;
    float txtz = tx / tz, tytz = ty / tz;
    e->xmul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
    e->ymul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
    float txc = fminf(limx, fmaxf(-limx, txtz)) * tz;
    float tyc = fminf(limy, fmaxf(-limy, tytz)) * tz;
    e->tx = txc; e->ty = tyc; e->tz = tz;
    float tz2 = tz * tz;
    e->J00 = fx / tz;
    e->J02 = -(fx * txc) / tz2;
    e->J11 = fy / tz;
    e->J12 = -(fy * tyc) / tz2;
    /* rows of the view rotation: R_r[k] = view[4k + r] */
    for (int k = 0; k < 3; ++k) {
        float R0 = view[4 * k + 0], R1 = view[4 * k + 1], R2 = view[4 * k + 2];
        e->m0[k] = fmaf(e->J02, R2, e->J00 * R0);
        e->m1[k] = fmaf(e->J12, R2, e->J11 * R1);
    }
    const float S00 = cov3D[0], S01 = cov3D[1], S02 = cov3D[2], S11 = cov3D[3], S12 = cov3D[4], S22 = cov3D[5];
    e->u0[0] = dot3f(S00, e->m0[0], S01, e->m0[1], S02, e->m0[2]);
    e->u0[1] = dot3f(S01, e->m0[0], S11, e->m0[1], S12, e->m0[2]);
    e->u0[2] = dot3f(S02, e->m0[0], S12, e->m0[1], S22, e->m0[2]);
    e->u1[0] = dot3f(S00, e->m1[0], S01, e->m1[1], S02, e->m1[2]);
    e->u1[1] = dot3f(S01, e->m1[0], S11, e->m1[1], S12, e->m1[2]);
    e->u1[2] = dot3f(S02, e->m1[0], S12, e->m1[1], S22, e->m1[2]);
    e->a = dot3f(e->m0[0], e->u0[0], e->m0[1], e->u0[1], e->m0[2], e->u0[2]) + 0.3f;
    e->b = dot3f(e->m0[0], e->u1[0], e->m0[1], e->u1[1], e->m0[2], e->u1[2]);
    e->c = dot3f(e->m1[0], e->u1[0], e->m1[1], e->u1[1], e->m1[2], e->u1[2]) + 0.3f;
}

static inline void sh_dir(const float *p, const float *campos, float d[3], float dorig[3]) {
    dorig[0] = p[0] - campos[0]; dorig[1] = p[1] - campos[1]; dorig[2] = p[2] - campos[2];
    float n = sqrtf(dot3f(dorig[0], dorig[0], dorig[1], dorig[1], dorig[2], dorig[2]));
    float inv = 1.f / n;
    d[0] = dorig[0] * inv; d[1] = dorig[1] * inv; d[2] = dorig[2] * inv;
}

/* A.1 step 9.  sh: [M][3] for this Gaussian. */
static void sh_to_rgb(int deg, const float *sh, const float dir[3], float rgb[3], uint8_t *clamped) {
    float x = dir[0], y = dir[1], z = dir[2];
    uint8_t cl = 0;
    for (int c = 0; c < 3; ++c) {
#define SHC(k) sh[(k) * 3 + c]
        float r;
        if (deg == 0) {
            r = fmaf(SH_C0, SHC(0), 0.5f);
        } else {
            r = SH_C0 * SHC(0);
            r = r - SH_C1 * y * SHC(1) + SH_C1 * z * SHC(2) - SH_C1 * x * SHC(3);
            if (deg > 1) {
                float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                r = r + SH_C2[0] * xy * SHC(4) + SH_C2[1] * yz * SHC(5) +
                    SH_C2[2] * (2.f * zz - xx - yy) * SHC(6) + SH_C2[3] * xz * SHC(7) +
                    SH_C2[4] * (xx - yy) * SHC(8);
                if (deg > 2) {
                    r = r + SH_C3[0] * y * (3.f * xx - yy) * SHC(9) + SH_C3[1] * xy * z * SHC(10) +
                        SH_C3[2] * y * (4.f * zz - xx - yy) * SHC(11) +
                        SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy) * SHC(12) +
                        SH_C3[4] * x * (4.f * zz - xx - yy) * SHC(13) +
                        SH_C3[5] * z * (xx - yy) * SHC(14) + SH_C3[6] * x * (xx - 3.f * yy) * SHC(15);
                }
            }
            r += 0.5f;
        }
#undef SHC
        if (r < 0.f) { cl |= (uint8_t)(1u << c); r = 0.f; }
        rgb[c] = r;
    }
    *clamped = cl;
}

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/*
 * A.1 -- preprocess forward.  Outputs are all per Gaussian; culled Gaussians keep radii=0, tiles_touched=0.
 *   rect: [P][4] = (min.x, min.y, max.x, max.y) in tiles.  cov3D_out: [P][6].  clamped: bit c set if channel c was clamped.
 *   shs may be NULL when colors_precomp is given; scales/rotations may be NULL when cov3D_precomp is given.
 */
void oracle_preprocess(int P, int D, int M, const float *means3D, const float *scales, const float *rotations,
                       const float *opacities, const float *shs, const float *colors_precomp,
                       const float *cov3D_precomp, float scale_modifier, const float *view, const float *proj,
                       const float *campos, int W, int H, float tanfovx, float tanfovy,
                       int32_t *radii, float *means2D, float *depths, float *cov3D_out, float *conic_opacity,
                       float *rgb, uint8_t *clamped, int32_t *rect, uint32_t *tiles_touched) {
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    const float fx = (float)W / (2.f * tanfovx), fy = (float)H / (2.f * tanfovy);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; ++i) {
        radii[i] = 0; tiles_touched[i] = 0;
        means2D[2 * i] = means2D[2 * i + 1] = 0.f; depths[i] = 0.f;
        for (int k = 0; k < 6; ++k) cov3D_out[6 * i + k] = 0.f;
        for (int k = 0; k < 4; ++k) { conic_opacity[4 * i + k] = 0.f; rect[4 * i + k] = 0; }
        rgb[3 * i] = rgb[3 * i + 1] = rgb[3 * i + 2] = 0.f; clamped[i] = 0;
        const float x = means3D[3 * i], y = means3D[3 * i + 1], z = means3D[3 * i + 2];
        float tx = dot3f(view[0], x, view[4], y, view[8], z) + view[12];
        float ty = dot3f(view[1], x, view[5], y, view[9], z) + view[13];
        float tz = dot3f(view[2], x, view[6], y, view[10], z) + view[14];
        if (tz <= 0.2f) continue;                                  /* near cull */
        float hx = dot3f(proj[0], x, proj[4], y, proj[8], z) + proj[12];
        float hy = dot3f(proj[1], x, proj[5], y, proj[9], z) + proj[13];
        float hw = dot3f(proj[3], x, proj[7], y, proj[11], z) + proj[15];
        float pw = 1.f / (hw + 0.0000001f);
        float px = hx * pw, py = hy * pw;
        float c3[6];
        if (cov3D_precomp) memcpy(c3, cov3D_precomp + 6 * i, sizeof c3);
        else cov3d_from_scale_rot(scales + 3 * i, scale_modifier, rotations + 4 * i, c3);
        Ewa e;
        ewa_project(tx, ty, tz, fx, fy, tanfovx, tanfovy, c3, view, &e);
        float det = fmaf(e.a, e.c, -(e.b * e.b));
        if (det == 0.f) continue;
        float det_inv = 1.f / det;
        float mid = 0.5f * (e.a + e.c);
        float sq = sqrtf(fmaxf(0.1f, fmaf(mid, mid, -det)));
        float lam = fmaxf(mid + sq, mid - sq);
        float rad = ceilf(3.f * sqrtf(lam));
        float pix_x = fmaf(px + 1.f, (float)W, -1.f) * 0.5f;
        float pix_y = fmaf(py + 1.f, (float)H, -1.f) * 0.5f;
        int irad = (int)rad;
        int rminx = clampi((int)((pix_x - (float)irad) / (float)TILE), 0, gx);
        int rminy = clampi((int)((pix_y - (float)irad) / (float)TILE), 0, gy);
        int rmaxx = clampi((int)((pix_x + (float)irad + (float)(TILE - 1)) / (float)TILE), 0, gx);
        int rmaxy = clampi((int)((pix_y + (float)irad + (float)(TILE - 1)) / (float)TILE), 0, gy);
        if ((rmaxx - rminx) * (rmaxy - rminy) == 0) continue;
        if (colors_precomp) {
            rgb[3 * i] = colors_precomp[3 * i]; rgb[3 * i + 1] = colors_precomp[3 * i + 1]; rgb[3 * i + 2] = colors_precomp[3 * i + 2];
        } else {
            float dir[3], dorig[3];
            sh_dir(means3D + 3 * i, campos, dir, dorig);
            sh_to_rgb(D, shs + (size_t)i * M * 3, dir, rgb + 3 * i, clamped + i);
        }
        depths[i] = tz;
        radii[i] = irad;
        means2D[2 * i] = pix_x; means2D[2 * i + 1] = pix_y;
        for (int k = 0; k < 6; ++k) cov3D_out[6 * i + k] = c3[k];
        conic_opacity[4 * i + 0] = e.c * det_inv;
        conic_opacity[4 * i + 1] = -e.b * det_inv;
        conic_opacity[4 * i + 2] = e.a * det_inv;
        conic_opacity[4 * i + 3] = opacities[i];
        rect[4 * i] = rminx; rect[4 * i + 1] = rminy; rect[4 * i + 2] = rmaxx; rect[4 * i + 3] = rmaxy;
        tiles_touched[i] = (uint32_t)((rmaxx - rminx) * (rmaxy - rminy));
    }
}

/* markVisible (K10): in-frustum test of the -w-pose fork = near cull only. */
void oracle_mark_visible(int P, const float *means3D, const float *view, uint8_t *present) {
    for (int i = 0; i < P; ++i) {
        const float x = means3D[3 * i], y = means3D[3 * i + 1], z = means3D[3 * i + 2];
        float tz = dot3f(view[2], x, view[6], y, view[10], z) + view[14];
        present[i] = tz > 0.2f;
    }
}

static int higher_msb(uint32_t n) { /* number of key bits needed for tile ids < n  (upstream getHigherMsb) */
    uint32_t msb = sizeof(n) * 4, step = msb;
    while (step > 1) {
        step /= 2;
        if (n >> msb) msb += step; else msb -= step;
    }
    if (n >> msb) msb++;
    return (int)msb;
}
int oracle_tile_bits(int W, int H) {
    return higher_msb((uint32_t)(((W + TILE - 1) / TILE) * ((H + TILE - 1) / TILE)));
}

/* A.2 -- total number of (tile, Gaussian) instances. */
int64_t oracle_count_instances(int P, const uint32_t *tiles_touched) {
    int64_t r = 0;
    for (int i = 0; i < P; ++i) r += tiles_touched[i];
    return r;
}

/*
 * A.2 -- duplicate with keys, stable LSD radix sort on bits [0, 32+tile_bits), tile ranges.
 *   keys_unsorted/keys_sorted: [R] u64 ; vals_*: [R] u32 ; ranges: [tiles][2] u32 (zero for untouched tiles).
 */
void oracle_bin(int P, int W, int H, const int32_t *radii, const float *depths, const int32_t *rect,
                const uint32_t *tiles_touched, int64_t R, uint64_t *keys_unsorted, uint32_t *vals_unsorted,
                uint64_t *keys_sorted, uint32_t *vals_sorted, uint32_t *ranges) {
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    int64_t off = 0;
    for (int i = 0; i < P; ++i) {
        if (radii[i] <= 0) continue;
        uint32_t dbits; memcpy(&dbits, depths + i, 4);
        for (int y = rect[4 * i + 1]; y < rect[4 * i + 3]; ++y)
            for (int x = rect[4 * i]; x < rect[4 * i + 2]; ++x) {
                uint64_t key = (uint64_t)(uint32_t)(y * gx + x);
                key = (key << 32) | dbits;
                keys_unsorted[off] = key; vals_unsorted[off] = (uint32_t)i; ++off;
            }
        (void)tiles_touched;
    }
    /* stable LSD radix sort, 16-bit digits, over the used bits only */
    const int bits = 32 + higher_msb((uint32_t)(gx * gy));
    uint64_t *ka = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(R ? R : 1));
    uint32_t *va = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(R ? R : 1));
    uint64_t *src_k = keys_sorted, *dst_k = ka; uint32_t *src_v = vals_sorted, *dst_v = va;
    memcpy(src_k, keys_unsorted, sizeof(uint64_t) * (size_t)R);
    memcpy(src_v, vals_unsorted, sizeof(uint32_t) * (size_t)R);
    size_t *cnt = (size_t *)malloc(sizeof(size_t) * 65537);
    for (int shift = 0; shift < bits; shift += 16) {
        memset(cnt, 0, sizeof(size_t) * 65537);
        for (int64_t r = 0; r < R; ++r) cnt[((src_k[r] >> shift) & 0xFFFF) + 1]++;
        for (int d = 0; d < 65536; ++d) cnt[d + 1] += cnt[d];
        for (int64_t r = 0; r < R; ++r) {
            size_t pos = cnt[(src_k[r] >> shift) & 0xFFFF]++;
            dst_k[pos] = src_k[r]; dst_v[pos] = src_v[r];
        }
        uint64_t *tk = src_k; src_k = dst_k; dst_k = tk;
        uint32_t *tv = src_v; src_v = dst_v; dst_v = tv;
    }
    if (src_k != keys_sorted) {
        memcpy(keys_sorted, src_k, sizeof(uint64_t) * (size_t)R);
        memcpy(vals_sorted, src_v, sizeof(uint32_t) * (size_t)R);
    }
    free(ka); free(va); free(cnt);
    memset(ranges, 0, sizeof(uint32_t) * 2 * (size_t)(gx * gy));
    for (int64_t r = 0; r < R; ++r) {
        uint32_t t = (uint32_t)(keys_sorted[r] >> 32);
        if (r == 0) ranges[2 * t] = 0;
        else {
            uint32_t tp = (uint32_t)(keys_sorted[r - 1] >> 32);
            if (t != tp) { ranges[2 * tp + 1] = (uint32_t)r; ranges[2 * t] = (uint32_t)r; }
        }
        if (r == R - 1) ranges[2 * t + 1] = (uint32_t)R;
    }
}

static inline float rel_margin(float v, float thr) {
    return fabsf(v - thr) / fmaxf(fabsf(thr), 1e-30f);
}

/*
 * A.3 -- blend forward.  out_color [3][H][W], out_depth/out_opacity/final_T [H][W], n_contrib [H][W] u32,
 * n_touched [P] i32 (zeroed here).  margin [H][W] (may be NULL): smallest relative distance of any compared
 * quantity to its threshold (alpha vs 1/255, T(1-alpha) vs 1e-4 and 0.5; power vs 0 uses |power|) seen while blending that pixel --
 * a checker uses it to set aside pixels whose discrete decisions are within float noise of flipping.
 */
void oracle_blend_forward(int P, int W, int H, const uint32_t *ranges, const uint32_t *point_list,
                          const float *means2D, const float *conic_opacity, const float *rgb, const float *depths,
                          const float *bg, float *out_color, float *out_depth, float *out_opacity, float *final_T,
                          uint32_t *n_contrib, int32_t *n_touched, float *margin) {
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    memset(n_touched, 0, sizeof(int32_t) * (size_t)P);
#pragma omp parallel for schedule(dynamic, 1) collapse(2)
    for (int ty = 0; ty < gy; ++ty)
        for (int tx = 0; tx < gx; ++tx) {
            const uint32_t r0 = ranges[2 * (ty * gx + tx)], r1 = ranges[2 * (ty * gx + tx) + 1];
            for (int ly = 0; ly < TILE; ++ly)
                for (int lx = 0; lx < TILE; ++lx) {
                    const int pxi = tx * TILE + lx, pyi = ty * TILE + ly;
                    if (pxi >= W || pyi >= H) continue;
                    const float pfx = (float)pxi, pfy = (float)pyi;
                    float T = 1.f, C0 = 0.f, C1 = 0.f, C2 = 0.f, Dp = 0.f, mg = 1e30f;
                    uint32_t contributor = 0, last = 0;
                    for (uint32_t r = r0; r < r1; ++r) {
                        const uint32_t g = point_list[r];
                        contributor++;
                        const float dx = means2D[2 * g] - pfx, dy = means2D[2 * g + 1] - pfy;
                        const float *co = conic_opacity + 4 * g;
                        const float power = -0.5f * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
                        if (margin && fabsf(power) < 1e-6f) mg = 0.f;
                        if (power > 0.f) continue;
                        const float a_raw = co[3] * expf(power);
                        const float alpha = fminf(0.99f, a_raw);
                        if (margin) mg = fminf(mg, rel_margin(a_raw, 1.f / 255.f));
                        if (alpha < 1.f / 255.f) continue;
                        const float test_T = T * (1.f - alpha);
                        if (margin) mg = fminf(mg, fminf(rel_margin(test_T, 0.0001f), rel_margin(test_T, 0.5f)));
                        if (test_T < 0.0001f) break;
                        const float w = alpha * T;
                        C0 += rgb[3 * g] * w; C1 += rgb[3 * g + 1] * w; C2 += rgb[3 * g + 2] * w;
                        Dp += depths[g] * w;
                        if (test_T > 0.5f) {
#pragma omp atomic
                            n_touched[g] += 1;
                        }
                        T = test_T;
                        last = contributor;
                    }
                    const size_t pix = (size_t)pyi * W + pxi;
                    final_T[pix] = T; n_contrib[pix] = last;
                    out_color[0 * (size_t)H * W + pix] = C0 + T * bg[0];
                    out_color[1 * (size_t)H * W + pix] = C1 + T * bg[1];
                    out_color[2 * (size_t)H * W + pix] = C2 + T * bg[2];
                    out_depth[pix] = Dp;
                    out_opacity[pix] = 1.f - T;
                    if (margin) margin[pix] = mg;
                }
        }
}

/*
 * A.4 -- blend backward.  Per-pair math in float32 exactly as A.4; the per-Gaussian sums are carried in
 * float64 (this is the checker: the CUDA path's float32 reduction-order noise sits around these values).
 * Accumulation order is deterministic: pixels row-major inside a tile, tiles merged in sorted-instance order.
 * Outputs (all zeroed here): dL_dmean2D [P][2] (NDC units: already multiplied by 0.5W / 0.5H),
 * dL_dconic [P][3] (x,y,w), dL_dopacity [P], dL_dcolor [P][3], dL_ddepth [P].
 * dL_dout_opacity may be NULL; it is used only when flags & FLAG_OPACITY_GRAD.
 */
void oracle_blend_backward(int P, int W, int H, int64_t R, const uint32_t *ranges, const uint32_t *point_list,
                           const float *means2D, const float *conic_opacity, const float *rgb, const float *depths,
                           const float *bg, const float *final_T, const uint32_t *n_contrib,
                           const float *dL_dout_color, const float *dL_dout_depth, const float *dL_dout_opacity,
                           int flags, float *dL_dmean2D, float *dL_dconic, float *dL_dopacity, float *dL_dcolor,
                           float *dL_ddepth) {
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    double *inst = (double *)calloc((size_t)(R ? R : 1) * 10, sizeof(double));
    const int use_og = (flags & FLAG_OPACITY_GRAD) && dL_dout_opacity;
#pragma omp parallel for schedule(dynamic, 1) collapse(2)
    for (int ty = 0; ty < gy; ++ty)
        for (int tx = 0; tx < gx; ++tx) {
            const uint32_t r0 = ranges[2 * (ty * gx + tx)], r1 = ranges[2 * (ty * gx + tx) + 1];
            for (int ly = 0; ly < TILE; ++ly)
                for (int lx = 0; lx < TILE; ++lx) {
                    const int pxi = tx * TILE + lx, pyi = ty * TILE + ly;
                    if (pxi >= W || pyi >= H) continue;
                    const size_t pix = (size_t)pyi * W + pxi;
                    const float pfx = (float)pxi, pfy = (float)pyi;
                    const float T_final = final_T[pix];
                    float T = T_final;
                    const uint32_t last = n_contrib[pix];
                    const float dpx[3] = {dL_dout_color[pix], dL_dout_color[(size_t)H * W + pix],
                                          dL_dout_color[2 * (size_t)H * W + pix]};
                    const float dpd = dL_dout_depth ? dL_dout_depth[pix] : 0.f;
                    float bg_dot = bg[0] * dpx[0] + bg[1] * dpx[1] + bg[2] * dpx[2];
                    if (use_og) bg_dot -= dL_dout_opacity[pix];   /* d(1-T_final)/dalpha = +T_final/(1-alpha) */
                    float acc[3] = {0.f, 0.f, 0.f}, acc_d = 0.f, last_alpha = 0.f, last_c[3] = {0.f, 0.f, 0.f}, last_d = 0.f;
                    for (uint32_t k = last; k-- > 0;) {           /* contributor index k (0-based) < n_contrib */
                        const uint32_t r = r0 + k;
                        if (r >= r1) continue;
                        const uint32_t g = point_list[r];
                        const float dx = means2D[2 * g] - pfx, dy = means2D[2 * g + 1] - pfy;
                        const float *co = conic_opacity + 4 * g;
                        const float power = -0.5f * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
                        if (power > 0.f) continue;
                        const float G = expf(power);
                        const float alpha = fminf(0.99f, co[3] * G);
                        if (alpha < 1.f / 255.f) continue;
                        T = T / (1.f - alpha);
                        const float w = alpha * T;
                        float dL_dalpha = 0.f;
                        double *o = inst + (size_t)r * 10;
                        for (int c = 0; c < 3; ++c) {
                            const float col = rgb[3 * g + c];
                            acc[c] = last_alpha * last_c[c] + (1.f - last_alpha) * acc[c];
                            last_c[c] = col;
                            dL_dalpha += (col - acc[c]) * dpx[c];
                            o[5 + c] += (double)(w * dpx[c]);
                        }
                        const float cd = depths[g];
                        acc_d = last_alpha * last_d + (1.f - last_alpha) * acc_d;
                        last_d = cd;
                        dL_dalpha += (cd - acc_d) * dpd;
                        o[8] += (double)(w * dpd);
                        dL_dalpha *= T;
                        last_alpha = alpha;
                        dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot;
                        const float dL_dG = co[3] * dL_dalpha;
                        const float gdx = G * dx, gdy = G * dy;
                        const float dG_ddelx = -gdx * co[0] - gdy * co[1];
                        const float dG_ddely = -gdy * co[2] - gdx * co[1];
                        o[0] += (double)(dL_dG * dG_ddelx * (0.5f * (float)W));
                        o[1] += (double)(dL_dG * dG_ddely * (0.5f * (float)H));
                        o[2] += (double)(-0.5f * gdx * dx * dL_dG);
                        o[3] += (double)(-0.5f * gdx * dy * dL_dG);
                        o[4] += (double)(-0.5f * gdy * dy * dL_dG);
                        o[9] += (double)(G * dL_dalpha);
                    }
                }
        }
    double *acc = (double *)calloc((size_t)(P ? P : 1) * 10, sizeof(double));
    for (int64_t r = 0; r < R; ++r) {
        const uint32_t g = point_list[r];
        for (int k = 0; k < 10; ++k) acc[(size_t)g * 10 + k] += inst[(size_t)r * 10 + k];
    }
    for (int i = 0; i < P; ++i) {
        const double *a = acc + (size_t)i * 10;
        dL_dmean2D[2 * i] = (float)a[0]; dL_dmean2D[2 * i + 1] = (float)a[1];
        dL_dconic[3 * i] = (float)a[2]; dL_dconic[3 * i + 1] = (float)a[3]; dL_dconic[3 * i + 2] = (float)a[4];
        dL_dcolor[3 * i] = (float)a[5]; dL_dcolor[3 * i + 1] = (float)a[6]; dL_dcolor[3 * i + 2] = (float)a[7];
        dL_ddepth[i] = (float)a[8];
        dL_dopacity[i] = (float)a[9];
    }
    free(inst); free(acc);
}

static inline void cross3(const float a[3], const float b[3], float o[3]) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}

/*
 * A.5 -- preprocess backward (cov2D backward + mean/depth/SH/cov3D backward + pose gradient).
 * Inputs: the blend-backward outputs.  Outputs (per Gaussian, zero for radii<=0):
 *   dL_dmeans3D [P][3], dL_dcov3D [P][6], dL_dsh [P][M][3] (NULL ok when colors_precomp),
 *   dL_dscales [P][3], dL_drots [P][4] (NULL ok when cov3D_precomp), dL_dtau [P][6] = (rho, theta).
 * tau convention: T_new = Exp(tau) T_w2c (utils/pose_utils.py:70-80) => d p_C / d rho = I, d p_C / d theta = -[p_C]x.
 */
void oracle_preprocess_backward(int P, int D, int M, const float *means3D, const int32_t *radii, const float *shs,
                                const uint8_t *clamped, const float *scales, const float *rotations,
                                const float *cov3D_precomp, float scale_modifier, const float *cov3D,
                                const float *view, const float *proj, const float *proj_raw, const float *campos,
                                int W, int H, float tanfovx, float tanfovy, int flags,
                                const float *dL_dmean2D, const float *dL_dconic, const float *dL_dcolor,
                                const float *dL_ddepth, float *dL_dmeans3D, float *dL_dcov3D, float *dL_dsh,
                                float *dL_dscales, float *dL_drots, float *dL_dtau) {
    const float fx = (float)W / (2.f * tanfovx), fy = (float)H / (2.f * tanfovy);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; ++i) {
        float dmean[3] = {0, 0, 0}, dcov[6] = {0, 0, 0, 0, 0, 0}, tau[6] = {0, 0, 0, 0, 0, 0};
        if (dL_dsh) memset(dL_dsh + (size_t)i * M * 3, 0, sizeof(float) * (size_t)M * 3);
        if (dL_dscales) dL_dscales[3 * i] = dL_dscales[3 * i + 1] = dL_dscales[3 * i + 2] = 0.f;
        if (dL_drots) dL_drots[4 * i] = dL_drots[4 * i + 1] = dL_drots[4 * i + 2] = dL_drots[4 * i + 3] = 0.f;
        if (radii[i] > 0) {
            const float x = means3D[3 * i], y = means3D[3 * i + 1], z = means3D[3 * i + 2];
            const float pc[3] = {dot3f(view[0], x, view[4], y, view[8], z) + view[12],
                                 dot3f(view[1], x, view[5], y, view[9], z) + view[13],
                                 dot3f(view[2], x, view[6], y, view[10], z) + view[14]};
            /* ---- A.5.1 cov2D backward ---- */
            Ewa e;
            ewa_project(pc[0], pc[1], pc[2], fx, fy, tanfovx, tanfovy, cov3D + 6 * i, view, &e);
            const float gxc = dL_dconic[3 * i], gyc = dL_dconic[3 * i + 1], gzc = dL_dconic[3 * i + 2];
            const float a = e.a, b = e.b, c = e.c;
            const float denom = a * c - b * b;
            const float k2 = 1.f / (denom * denom + 0.0000001f);
            float da = 0.f, db = 0.f, dc = 0.f;
            if (k2 != 0.f) {
                da = k2 * (-c * c * gxc + 2.f * b * c * gyc + (denom - a * c) * gzc);
                dc = k2 * (-a * a * gzc + 2.f * a * b * gyc + (denom - a * c) * gxc);
                db = k2 * 2.f * (b * c * gxc - (denom + 2.f * b * b) * gyc + a * b * gzc);
                const float *m0 = e.m0, *m1 = e.m1;
                dcov[0] = m0[0] * m0[0] * da + m0[0] * m1[0] * db + m1[0] * m1[0] * dc;
                dcov[3] = m0[1] * m0[1] * da + m0[1] * m1[1] * db + m1[1] * m1[1] * dc;
                dcov[5] = m0[2] * m0[2] * da + m0[2] * m1[2] * db + m1[2] * m1[2] * dc;
                dcov[1] = 2.f * m0[0] * m0[1] * da + (m0[0] * m1[1] + m0[1] * m1[0]) * db + 2.f * m1[0] * m1[1] * dc;
                dcov[2] = 2.f * m0[0] * m0[2] * da + (m0[0] * m1[2] + m0[2] * m1[0]) * db + 2.f * m1[0] * m1[2] * dc;
                dcov[4] = 2.f * m0[2] * m0[1] * da + (m0[1] * m1[2] + m0[2] * m1[1]) * db + 2.f * m1[1] * m1[2] * dc;
            }
            float dm0[3], dm1[3];
            for (int k = 0; k < 3; ++k) {
                dm0[k] = 2.f * e.u0[k] * da + e.u1[k] * db;
                dm1[k] = 2.f * e.u1[k] * dc + e.u0[k] * db;
            }
            float R0[3], R1[3], R2[3];
            for (int k = 0; k < 3; ++k) { R0[k] = view[4 * k]; R1[k] = view[4 * k + 1]; R2[k] = view[4 * k + 2]; }
            const float dJ00 = R0[0] * dm0[0] + R0[1] * dm0[1] + R0[2] * dm0[2];
            const float dJ02 = R2[0] * dm0[0] + R2[1] * dm0[1] + R2[2] * dm0[2];
            const float dJ11 = R1[0] * dm1[0] + R1[1] * dm1[1] + R1[2] * dm1[2];
            const float dJ12 = R2[0] * dm1[0] + R2[1] * dm1[1] + R2[2] * dm1[2];
            const float tzi = 1.f / e.tz, tz2 = tzi * tzi, tz3 = tz2 * tzi;
            float dt[3];
            dt[0] = e.xmul * -fx * tz2 * dJ02;
            dt[1] = e.ymul * -fy * tz2 * dJ12;
            dt[2] = -fx * tz2 * dJ00 - fy * tz2 * dJ11 + (2.f * fx * e.tx) * tz3 * dJ02 + (2.f * fy * e.ty) * tz3 * dJ12;
            for (int k = 0; k < 3; ++k) dmean[k] = R0[k] * dt[0] + R1[k] * dt[1] + R2[k] * dt[2];
            /* pose through t (uses the clamped t, as the recomputed forward does) */
            {
                const float tcl[3] = {e.tx, e.ty, e.tz};
                float cr[3];
                cross3(tcl, dt, cr);
                for (int k = 0; k < 3; ++k) { tau[k] += dt[k]; tau[3 + k] += cr[k]; }
            }
            /* pose through the view rotation: dL/dR[r][k], columns r_k = (R0[k],R1[k],R2[k]) */
            for (int k = 0; k < 3; ++k) {
                const float col[3] = {R0[k], R1[k], R2[k]};
                const float g[3] = {e.J00 * dm0[k], e.J11 * dm1[k], e.J02 * dm0[k] + e.J12 * dm1[k]};
                float cr[3];
                cross3(col, g, cr);
                tau[3] += cr[0]; tau[4] += cr[1]; tau[5] += cr[2];
            }
            /* ---- A.5.2 mean2D / depth ---- */
            const float hx = dot3f(proj[0], x, proj[4], y, proj[8], z) + proj[12];
            const float hy = dot3f(proj[1], x, proj[5], y, proj[9], z) + proj[13];
            const float hw = dot3f(proj[3], x, proj[7], y, proj[11], z) + proj[15];
            const float mw = 1.f / (hw + 0.0000001f);
            const float mul1 = hx * mw * mw, mul2 = hy * mw * mw;
            const float g2x = dL_dmean2D[2 * i], g2y = dL_dmean2D[2 * i + 1];
            dmean[0] += (proj[0] * mw - proj[3] * mul1) * g2x + (proj[1] * mw - proj[3] * mul2) * g2y;
            dmean[1] += (proj[4] * mw - proj[7] * mul1) * g2x + (proj[5] * mw - proj[7] * mul2) * g2y;
            dmean[2] += (proj[8] * mw - proj[11] * mul1) * g2x + (proj[9] * mw - proj[11] * mul2) * g2y;
            {
                const float al = mw, be = -hx * mw * mw, ga = -hy * mw * mw;
                const float pa = proj_raw[0], pb = proj_raw[5], pe = proj_raw[11];
                float d1[3] = {al * pa, 0.f, be * pe}, d2[3] = {0.f, al * pb, ga * pe};
                if (flags & FLAG_EXACT_PP) { d1[2] += al * proj_raw[8]; d2[2] += al * proj_raw[9]; }
                float v[3] = {g2x * d1[0] + g2y * d2[0], g2x * d1[1] + g2y * d2[1], g2x * d1[2] + g2y * d2[2]};
                float cr[3];
                cross3(pc, v, cr);
                for (int k = 0; k < 3; ++k) { tau[k] += v[k]; tau[3 + k] += cr[k]; }
            }
            {
                const float dz = dL_ddepth[i];
                dmean[0] += dz * view[2]; dmean[1] += dz * view[6]; dmean[2] += dz * view[10];
                tau[2] += dz;
                tau[3] += dz * pc[1];
                tau[4] += dz * -pc[0];
            }
            /* ---- SH backward ---- */
            if (dL_dsh && shs) {
                float dir[3], dorig[3];
                sh_dir(means3D + 3 * i, campos, dir, dorig);
                const float *sh = shs + (size_t)i * M * 3;
                float *dsh = dL_dsh + (size_t)i * M * 3;
                float dRGB[3];
                for (int ch = 0; ch < 3; ++ch) dRGB[ch] = dL_dcolor[3 * i + ch] * ((clamped[i] >> ch) & 1 ? 0.f : 1.f);
                const float sx = dir[0], sy = dir[1], sz = dir[2];
                float ddir[3] = {0, 0, 0};
                for (int ch = 0; ch < 3; ++ch) {
#define SHC(k) sh[(k) * 3 + ch]
#define DSH(k) dsh[(k) * 3 + ch]
                    const float g = dRGB[ch];
                    float ddx = 0.f, ddy = 0.f, ddz = 0.f;
                    DSH(0) = SH_C0 * g;
                    if (D > 0) {
                        DSH(1) = -SH_C1 * sy * g; DSH(2) = SH_C1 * sz * g; DSH(3) = -SH_C1 * sx * g;
                        ddx = -SH_C1 * SHC(3); ddy = -SH_C1 * SHC(1); ddz = SH_C1 * SHC(2);
                        if (D > 1) {
                            const float xx = sx * sx, yy = sy * sy, zz = sz * sz, xy = sx * sy, yz = sy * sz, xz = sx * sz;
                            DSH(4) = SH_C2[0] * xy * g; DSH(5) = SH_C2[1] * yz * g;
                            DSH(6) = SH_C2[2] * (2.f * zz - xx - yy) * g; DSH(7) = SH_C2[3] * xz * g;
                            DSH(8) = SH_C2[4] * (xx - yy) * g;
                            ddx += SH_C2[0] * sy * SHC(4) + SH_C2[2] * 2.f * -sx * SHC(6) + SH_C2[3] * sz * SHC(7) + SH_C2[4] * 2.f * sx * SHC(8);
                            ddy += SH_C2[0] * sx * SHC(4) + SH_C2[1] * sz * SHC(5) + SH_C2[2] * 2.f * -sy * SHC(6) + SH_C2[4] * 2.f * -sy * SHC(8);
                            ddz += SH_C2[1] * sy * SHC(5) + SH_C2[2] * 2.f * 2.f * sz * SHC(6) + SH_C2[3] * sx * SHC(7);
                            if (D > 2) {
                                DSH(9) = SH_C3[0] * sy * (3.f * xx - yy) * g;
                                DSH(10) = SH_C3[1] * xy * sz * g;
                                DSH(11) = SH_C3[2] * sy * (4.f * zz - xx - yy) * g;
                                DSH(12) = SH_C3[3] * sz * (2.f * zz - 3.f * xx - 3.f * yy) * g;
                                DSH(13) = SH_C3[4] * sx * (4.f * zz - xx - yy) * g;
                                DSH(14) = SH_C3[5] * sz * (xx - yy) * g;
                                DSH(15) = SH_C3[6] * sx * (xx - 3.f * yy) * g;
                                ddx += SH_C3[0] * SHC(9) * 3.f * 2.f * xy + SH_C3[1] * SHC(10) * yz + SH_C3[2] * SHC(11) * -2.f * xy +
                                       SH_C3[3] * SHC(12) * -3.f * 2.f * xz + SH_C3[4] * SHC(13) * (-3.f * xx + 4.f * zz - yy) +
                                       SH_C3[5] * SHC(14) * 2.f * xz + SH_C3[6] * SHC(15) * 3.f * (xx - yy);
                                ddy += SH_C3[0] * SHC(9) * 3.f * (xx - yy) + SH_C3[1] * SHC(10) * xz +
                                       SH_C3[2] * SHC(11) * (-3.f * yy + 4.f * zz - xx) + SH_C3[3] * SHC(12) * -3.f * 2.f * yz +
                                       SH_C3[4] * SHC(13) * -2.f * xy + SH_C3[5] * SHC(14) * -2.f * yz + SH_C3[6] * SHC(15) * -3.f * 2.f * xy;
                                ddz += SH_C3[1] * SHC(10) * xy + SH_C3[2] * SHC(11) * 4.f * 2.f * yz +
                                       SH_C3[3] * SHC(12) * 3.f * (2.f * zz - xx - yy) + SH_C3[4] * SHC(13) * 4.f * 2.f * xz +
                                       SH_C3[5] * SHC(14) * (xx - yy);
                            }
                        }
                    }
                    ddir[0] += ddx * g; ddir[1] += ddy * g; ddir[2] += ddz * g;
#undef SHC
#undef DSH
                }
                if (D > 0) {
                    /* through dir = normalize(p - campos) */
                    const float s2 = dorig[0] * dorig[0] + dorig[1] * dorig[1] + dorig[2] * dorig[2];
                    const float inv32 = 1.f / sqrtf(s2 * s2 * s2);
                    float dm[3];
                    dm[0] = ((s2 - dorig[0] * dorig[0]) * ddir[0] - dorig[1] * dorig[0] * ddir[1] - dorig[2] * dorig[0] * ddir[2]) * inv32;
                    dm[1] = (-dorig[0] * dorig[1] * ddir[0] + (s2 - dorig[1] * dorig[1]) * ddir[1] - dorig[2] * dorig[1] * ddir[2]) * inv32;
                    dm[2] = (-dorig[0] * dorig[2] * ddir[0] - dorig[1] * dorig[2] * ddir[1] + (s2 - dorig[2] * dorig[2]) * ddir[2]) * inv32;
                    for (int k = 0; k < 3; ++k) { dmean[k] += dm[k]; tau[k] += -dm[k]; }   /* A.5.2 / A.6 item 7 */
                }
            }
            /* ---- cov3D -> scale / rotation ---- */
            if (!cov3D_precomp && dL_dscales && dL_drots) {
                const float *q = rotations + 4 * i;
                float Rq[9];
                quat_to_R(q, Rq);
                const float s[3] = {scale_modifier * scales[3 * i], scale_modifier * scales[3 * i + 1], scale_modifier * scales[3 * i + 2]};
                /* G = dL/dSigma as a full symmetric matrix (off-diagonals halved) */
                const float Gm[9] = {dcov[0], 0.5f * dcov[1], 0.5f * dcov[2], 0.5f * dcov[1], dcov[3], 0.5f * dcov[4],
                                     0.5f * dcov[2], 0.5f * dcov[4], dcov[5]};
                /* A = Rq diag(s); dL/dA = 2 G A */
                float dA[9];
                for (int r = 0; r < 3; ++r)
                    for (int k = 0; k < 3; ++k) {
                        float acc = 0.f;
                        for (int j = 0; j < 3; ++j) acc += Gm[r * 3 + j] * (Rq[j * 3 + k] * s[k]);
                        dA[r * 3 + k] = 2.f * acc;
                    }
                float Q[9];
                for (int k = 0; k < 3; ++k) {
                    float ds = 0.f;
                    for (int r = 0; r < 3; ++r) { ds += dA[r * 3 + k] * Rq[r * 3 + k]; Q[r * 3 + k] = dA[r * 3 + k] * s[k]; }
                    dL_dscales[3 * i + k] = ds;           /* upstream: no extra scale_modifier factor */
                }
                const float r = q[0], qx = q[1], qy = q[2], qz = q[3];
                dL_drots[4 * i + 0] = 2.f * (-qz * Q[1] + qy * Q[2] + qz * Q[3] - qx * Q[5] - qy * Q[6] + qx * Q[7]);
                dL_drots[4 * i + 1] = 2.f * (qy * Q[1] + qz * Q[2] + qy * Q[3] - 2.f * qx * Q[4] - r * Q[5] + qz * Q[6] + r * Q[7] - 2.f * qx * Q[8]);
                dL_drots[4 * i + 2] = 2.f * (-2.f * qy * Q[0] + qx * Q[1] + r * Q[2] + qx * Q[3] + qz * Q[5] - r * Q[6] + qz * Q[7] - 2.f * qy * Q[8]);
                dL_drots[4 * i + 3] = 2.f * (-2.f * qz * Q[0] - r * Q[1] + qx * Q[2] + r * Q[3] - 2.f * qz * Q[4] + qy * Q[5] + qx * Q[6] + qy * Q[7]);
            }
        }
        for (int k = 0; k < 3; ++k) dL_dmeans3D[3 * i + k] = dmean[k];
        for (int k = 0; k < 6; ++k) { dL_dcov3D[6 * i + k] = dcov[k]; dL_dtau[6 * i + k] = tau[k]; }
    }
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm of bench.py asks for the host's cores explicitly */
void oracle_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
