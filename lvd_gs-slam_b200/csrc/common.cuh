// common.cuh -- shared definitions for the sm_100a rasterizer kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/lvdgs.h"

namespace lvdgs {

constexpr int TILE = LVDGS_TILE;
constexpr int TILE_PIX = TILE * TILE;

// ---- error plumbing (api.cu) ----
void set_error(const char *fmt, ...);
void count_launch();
int profile_mark(const char *name, cudaStream_t s);   // profiling on: event after the launch (or a host-side marker)
int profile_pre(cudaStream_t s);                      // profiling on: event right before the launch
extern thread_local int g_debug_sync;

#define LVDGS_CHECK(expr)                                                                        \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            lvdgs::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,               \
                             cudaGetErrorString(_e));                                            \
            return 1;                                                                            \
        }                                                                                        \
    } while (0)

// after every kernel launch: count it, catch launch errors, optionally synchronise (debug)
#define LVDGS_PRE(stream)                                                                        \
    do {                                                                                         \
        if (lvdgs::profile_pre(stream)) return 1;                                                \
    } while (0)

#define LVDGS_LAUNCHED(stream, name)                                                             \
    do {                                                                                         \
        lvdgs::count_launch();                                                                   \
        LVDGS_CHECK(cudaGetLastError());                                                         \
        if (lvdgs::profile_mark(name, stream)) return 1;                                         \
        if (lvdgs::g_debug_sync) LVDGS_CHECK(cudaStreamSynchronize(stream));                     \
    } while (0)

// ---- programmatic dependent launch (sm_90+): a kernel launched with launch_after_kernel() may be scheduled while the
// kernel before it in the stream is still draining; it must execute pdl_wait() before it touches anything that kernel
// wrote (everything older in the stream is complete by then).  No kernel here triggers its dependents early, so the only
// effect is that launch latency and block scheduling overlap the predecessor's tail.  LVDGS_PDL=0 (environment) or an
// active profiler (events between the launches) falls back to plain stream order.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
bool pdl_enabled();                                    // api.cu
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_after_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// Per-device state: function attributes (opt-in shared memory sizes) and device properties belong to ONE device; a process
// that drives several GPUs (one engine per device, or the tests' two-GPU runs) needs them once per device, not once per
// process.  current_device() is the device the calling thread has selected (lvdgs_set_device / cudaSetDevice).
constexpr int MAX_DEVICES = 64;
static inline int current_device() { int d = 0; cudaGetDevice(&d); return d < 0 || d >= MAX_DEVICES ? 0 : d; }

static inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }
static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---- canonical arithmetic: explicit IEEE ops, never re-contracted by nvcc ----
__device__ __forceinline__ float dot3c(float a0, float b0, float a1, float b1, float a2, float b2) {
    return __fmaf_rn(a2, b2, __fmaf_rn(a1, b1, __fmul_rn(a0, b0)));
}

// ---- shared-memory loads through 32-bit shared-window addresses: one IMAD + LDS in the blend hot loops (indexing a
// __shared__ array from divergent code makes nvcc 12.9 rebuild the cluster-window base, S2R CgaCtaId + LEA, per access)
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    uint32_t a;   // volatile: computed once where written, never rematerialised inside the loops
    asm volatile("{ .reg .u64 t; cvta.to.shared.u64 t, %1; cvt.u32.u64 %0, t; }" : "=r"(a) : "l"(p));
    return a;
}
__device__ __forceinline__ float4 lds128(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ float2 lds64(uint32_t a) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
// 2^x on the MUFU pipe (flushes to zero below 2^-126, which the alpha >= 1/255 test discards anyway)
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// 1/x on the MUFU pipe for normal-range x (no scaling around denormals, unlike __fdividef / __frcp_rn)
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// ---- packed FP32 pairs (sm_100a FFMA2 / FMUL2 / FADD2: two IEEE fp32 operations per issue slot).  A `bc(x)` operand
// (both halves equal) costs nothing: ptxas folds it into the instruction's scalar-broadcast operand form.
// Loop-carried state is kept as two scalar floats and packed at the use (pk() of two live floats is free): a 64-bit
// value carried around a loop back-edge costs two IMAD.MOV per iteration with ptxas 12.9 (8 of the 58 instructions of
// the forward blend's visit, 4 of the backward's, before this was changed).
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ f32x2 bc(float x) { return pk(x, x); }
__device__ __forceinline__ float lo_of(f32x2 v) { return __uint_as_float((uint32_t)v); }
__device__ __forceinline__ float hi_of(f32x2 v) { return __uint_as_float((uint32_t)(v >> 32)); }
__device__ __forceinline__ float hsum(f32x2 v) { return lo_of(v) + hi_of(v); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
constexpr float LOG2E = 1.4426950408889634f;
// One staged instance of a blend batch: the three loads of the hot loops share ONE address computation (immediate
// offsets 0 / 16 / 32), and 48-byte records keep the staging stores (one record per lane) bank-conflict free.
struct __align__(16) BlendRec {
    float2 xy;          // pixel-space mean
    uint32_t id;        // Gaussian index
    uint32_t pad;
    float4 co;          // conic prescaled to base 2 (-0.5 log2e A, -log2e B, -0.5 log2e C) + opacity
    float4 cd;          // rgb + depth
};
static_assert(sizeof(BlendRec) == 48, "BlendRec layout");
// pixel x coordinate assigned to a finished / out-of-image pixel: every Gaussian then evaluates to alpha = 0, so the
// hot loop needs no per-pixel "done" test (dx^2 ~ 1e36 stays finite in float)
constexpr float PIX_PARKED = 1.0e18f;

// Exact test "does the ellipse {d : A dx^2 + 2B dx dy + C dy^2 <= lvl} around (mx,my) reach the rectangle of pixel centres
// [X0,X1] x [Y0,Y1]?" -- the minimum of the convex quadratic over the rectangle is 0 if the centre is inside, else it
// lies on one of the four edges, where it is a clamped 1-D parabola minimum.  Used by the blend staging on the pixel
// blocks that survive the bounding-box test; `lvl` carries the same 1% + 0.02 slack as the box, so the test stays
// conservative with respect to the reference's alpha >= 1/255 decision.
__device__ __forceinline__ bool ellipse_reaches_rect(float mx, float my, float A, float B, float C, float rA, float rC,
                                                     float lvl, float X0, float Y0, float X1, float Y1) {
    const float dx0 = X0 - mx, dx1 = X1 - mx, dy0 = Y0 - my, dy1 = Y1 - my;
    if (dx0 <= 0.f && dx1 >= 0.f && dy0 <= 0.f && dy1 >= 0.f) return true;
    float q = 3.0e38f;
    {
        const float t0 = fminf(fmaxf(-B * dx0 * rC, dy0), dy1), t1 = fminf(fmaxf(-B * dx1 * rC, dy0), dy1);
        q = fminf(q, A * dx0 * dx0 + (2.f * B * dx0 + C * t0) * t0);
        q = fminf(q, A * dx1 * dx1 + (2.f * B * dx1 + C * t1) * t1);
    }
    {
        const float t0 = fminf(fmaxf(-B * dy0 * rA, dx0), dx1), t1 = fminf(fmaxf(-B * dy1 * rA, dx0), dx1);
        q = fminf(q, C * dy0 * dy0 + (2.f * B * dy0 + A * t0) * t0);
        q = fminf(q, C * dy1 * dy1 + (2.f * B * dy1 + A * t1) * t1);
    }
    return !(q > lvl);        // NaN -> keep
}

struct CameraConst {   // staged once per block in shared memory
    float view[16];
    float proj[16];
    float campos[3];
};

// number of key bits that cover tile ids < n (same rule as upstream's getHigherMsb)
static inline int tile_bits(uint32_t n) {
    uint32_t msb = sizeof(n) * 4, step = msb;
    while (step > 1) {
        step /= 2;
        if (n >> msb) msb += step; else msb -= step;
    }
    if (n >> msb) msb++;
    return (int)msb;
}

// ---- kernel launchers (one per .cu) ----
struct GeomPtrs {
    float *depths; float4 *means2D; float4 *conic_opacity; float4 *rgbd; short4 *rect;
    uint32_t *tiles_touched; uint32_t *point_offsets; uint8_t *clamped;
    uint32_t *block_sums;             // [ceil(P/256)] instances per preprocess block -> exclusive offsets (binning_prep)
    uint32_t *num_instances;          // [0] R, [1] longest tile list (binning_prep), [2] visible Gaussians (key emission)
    uint32_t *visible_list;           // [P] indices of the Gaussians with tiles_touched > 0 (written by the key emission, in
                                      // block order): the preprocess backward of the accumulate / pose-only modes walks it
};
struct BinPtrs {
    uint64_t *keys[2]; uint32_t *vals[2]; void *sort_ws; int32_t *sorted_sel;
};
struct ImgPtrs {
    float *final_T; uint32_t *n_contrib; uint2 *ranges;
    int32_t *tile_grid;               // [(gy+1)*(gx+1)] 2-D difference array of the tile rects -> per-tile instance counts
    uint32_t *tile_order;             // [tiles] tile ids, heaviest instance lists first (launch order of the blend kernels)
    uint32_t *sort_hist;              // [8][256] digit histograms of the (tile|depth) keys, exclusive-scanned by binning_prep
    uint32_t *tile_cursor;            // [tiles][CURSOR_STRIDE] instances emitted so far into each tile's segment (tile-segment sort)
};
// one cursor per 32-byte sector: same-sector atomics serialise in L2, and all cursors packed would be ~60 lines
#ifndef LVDGS_CURSOR_STRIDE
#define LVDGS_CURSOR_STRIDE 8
#endif
constexpr int CURSOR_STRIDE = LVDGS_CURSOR_STRIDE;
constexpr int SORT_MAX_PASSES = 8;
constexpr int SORT_BINS = 256;

int launch_preprocess_forward(const lvdgs_raster_params &p, const float *means3D, const float *colors_precomp,
                              const float *opacities, const float *scales, const float *rotations,
                              const float *cov3D_precomp, const float *view, const float *proj, const float *shs,
                              const float *campos, int32_t *radii, int32_t *n_touched, const GeomPtrs &g, const ImgPtrs &im, cudaStream_t s);
int launch_binning_prep(int P, int W, int H, int end_bit, const GeomPtrs &g, const ImgPtrs &im, uint32_t *host_rb, uint32_t host_seq, cudaStream_t s);
int launch_emit_keys(int P, int W, int H, const GeomPtrs &g, int64_t capacity, uint64_t *keys, uint32_t *vals,
                     uint32_t *tile_cursor, const uint2 *ranges, int32_t *sel_out, cudaStream_t s);
// tile-segment sort (tile_sort.cu): seg holds each tile's (depth bits << 32 | Gaussian) words in ranges[tile], unordered
int launch_tile_sort(int tiles, int64_t capacity, const uint32_t *n_dev, const uint2 *ranges, const uint32_t *tile_order,
                     uint64_t *seg, uint64_t *keys_out, uint32_t *vals_out, bool long_lists, cudaStream_t s);
int tile_sort_long_threshold();

size_t sort_workspace_bytes(int64_t n);
// pre_hist: optional [passes][256] exclusive-scanned digit histograms (device); when given, the histogram pass over the
// keys is skipped (the forward derives them in preprocess / binning_prep).
int launch_sort_pairs(int64_t n, const uint32_t *n_dev, uint64_t *keys0, uint64_t *keys1, uint32_t *vals0,
                      uint32_t *vals1, int end_bit, void *ws, size_t ws_bytes, const uint32_t *pre_hist, int *selector,
                      cudaStream_t s);

int launch_blend_forward(int W, int H, int64_t capacity, const uint32_t *n_dev, const uint2 *ranges, const uint32_t *point_list, const GeomPtrs &g,
                         const uint32_t *tile_order, const float *bg, float *out_color, float *out_depth,
                         float *out_opacity, float *final_T, uint32_t *n_contrib, int32_t *n_touched, cudaStream_t s);

// Per-Gaussian accumulators produced by the blend backward: one 48-byte row per Gaussian so that a warp can
// commit its partial sums with three 128-bit vector reductions (red.global.add.v4.f32).
// The geometric slots are the moments of m = G * dL/dalpha over the Gaussian's pixels (d = mean2D - pixel):
//   [0] S_x = sum m dx   [1] S_y = sum m dy   [2] S_xx = sum m dx^2   [3] S_xy = sum m dx dy
//   [4] S_yy = sum m dy^2   [5] S_0 = sum m (= dL_dopacity)   [6] dL_ddepth   [7] -
//   [8..10] dL_dcolor rgb   [11] -
// from which the preprocess backward forms, with (A,B,C,o) = conic_opacity:
//   dL_dmean2D = -o (A S_x + B S_y, B S_x + C S_y) * (W/2, H/2),  dL_dconic = -o/2 (S_xx, S_xy, S_yy).
constexpr int ACC_STRIDE = 12;
struct BlendGradPtrs {
    float *acc;          // [P][ACC_STRIDE], zeroed by the API before the blend backward
};
int launch_blend_backward(int P, int W, int H, int64_t R, const uint2 *ranges, const uint32_t *point_list,
                          const uint32_t *tile_order, const GeomPtrs &g, const float *bg, const float *final_T, const uint32_t *n_contrib,
                          const float *dL_dout_color, const float *dL_dout_depth, const float *dL_dout_opacity,
                          int flags, bool moments_only, const BlendGradPtrs &o, float *zero6, cudaStream_t s);

int launch_preprocess_backward(const lvdgs_raster_params &p, const float *means3D, const int32_t *radii,
                               const float *shs, const float *scales, const float *rotations,
                               const float *cov3D_precomp, const float *view, const float *proj,
                               const float *proj_raw, const float *campos, const GeomPtrs &g,
                               const BlendGradPtrs &bgp, bool colors_are_precomp, float *dL_dmeans2D,
                               float *dL_dcolors, float *dL_dopacity, float *dL_dmeans3D, float *dL_dcov3D,
                               float *dL_dsh, float *dL_dscales, float *dL_drots, float *dL_dtau,
                               float *dL_dtau_sum, bool tau_sum_zeroed, cudaStream_t s);     // uses g.visible_list when no output needs the culled rows

int launch_mark_visible(int P, const float *means3D, const float *view, uint8_t *present, cudaStream_t s);

int launch_adam_step(int64_t n, float *params, const float *grads, float *exp_avg, float *exp_avg_sq, int groups,
                     const int64_t *group_end, const float *lr, double beta1, double beta2, double eps, int step,
                     cudaStream_t s);

size_t fused_loss_workspace_bytes();
int launch_fused_loss(int W, int H, const float *color, const float *depth, const float *opacity, const float *gt_color,
                      const float *gt_depth, const float *grad_mask, const float *exposure, float thr, float w_rgb,
                      float w_depth, int flags, float *g_color, float *g_depth, float *g_opacity, float *out, void *ws,
                      size_t ws_bytes, cudaStream_t s);
size_t masked_ssim_workspace_bytes(int W, int H);
int launch_masked_ssim_loss(int W, int H, const float *image, const float *gt, const uint8_t *mask, const float *bg,
                            const float *depth, const float *mono, float lambda_dssim, float depth_lambda, float *g_image,
                            float *g_depth, float *out, void *ws, size_t ws_bytes, cudaStream_t s);
int launch_covis(int64_t n, const void *a, const void *b, int elem, unsigned long long *out, cudaStream_t s);
int launch_n_obs(int64_t n, int K, const void *const *masks_dev, int elem, int32_t *n_obs, cudaStream_t s);
size_t compact_workspace_bytes(int64_t n);
int launch_compact_count(int64_t n, const uint8_t *keep, void *ws, size_t ws_bytes, uint32_t **count_dev, cudaStream_t s);
int launch_compact_move(int64_t n, const uint8_t *keep, const void *ws, int n_arrays, const float *const *src,
                        float *const *dst, const int32_t *widths, cudaStream_t s);

int launch_gather_rows(int64_t n_idx, const int64_t *idx, int64_t n_src_rows, int n_arrays, const float *const *src,
                       float *const *dst, const int32_t *widths, cudaStream_t s);
int launch_gaussian_activate(int64_t P, const float *raw_o, const float *raw_s, const float *raw_q, float *o, float *sc, float *q,
                             cudaStream_t s);
int launch_gaussian_activation_backward(int64_t P, const float *o, const float *sc, const float *q, const float *raw_q, float *g_o,
                                        float *g_s, float *g_q, cudaStream_t s);
int launch_pose_step(lvdgs_pose_state *state, const float *g_tau, const float *g_exposure, float lr_rot, float lr_trans,
                     float lr_exp, double beta1, double beta2, double eps, int step, float threshold, cudaStream_t s);

int launch_exchange_adam(int world, int rank, const float *const *grad_ptrs, float *const *param_ptrs, float *const *act_ptrs,
                         int64_t lo, int64_t hi, float *exp_avg, float *exp_avg_sq, int groups, const int64_t *group_end,
                         const float *lr, const int64_t *act_offsets, int64_t act_total, double beta1, double beta2, double eps,
                         int step, const float *mc_grad, float *mc_param, float *mc_act, int act_mode, cudaStream_t s);

int launch_fp32_peak(int blocks, int iters, int mode, float *out, double *fmas, cudaStream_t s);

size_t dist2_workspace_bytes(int P);
int launch_dist2(int P, const float *points, float *mean_dists, void *ws, size_t ws_bytes, cudaStream_t s);

}  // namespace lvdgs
