// slam_ops.cu -- the callers either side of the rasterizer: the photometric +
// depth losses that produce the rasterizer's upstream gradients, the keyframe-covisibility counts on n_touched masks,
// and the row compaction behind prune_points.  All HBM-bound single passes; each replaces a chain of torch elementwise
// kernels and their full-size temporaries in the reference.
#include "common.cuh"
#include <cmath>

namespace lvdgs {

// ---------------------------------------------------------------------------------------------------------
// N3: fused tracking / mapping loss (utils/slam_utils.py:42-121) with its gradient, in one pass over the pixels.
//   loss = w_rgb * mean_{c,p} o(p) | m(p) (ea I_c(p) + eb) - m(p) gt_c(p) |  +  w_d * mean_p | md(p) D(p) - md(p) gtD(p) |
//   ea = exp(exposure_a), eb = exposure_b (image_ab of get_loss_tracking / get_loss_mapping),
//   m  = (sum_c gt_c > rgb_boundary_threshold) * grad_mask          (grad_mask only in the tracking loss, :59)
//   o  = rendered opacity (tracking, :60) or 1 (mapping)
//   md = (gtD > 0.01) * (opacity > 0.95 in the tracking rgbd loss, :73-74)
// Outputs: dL/dI [3,H,W], dL/dD [H,W], dL/dopacity [H,W] (optional), and out[4] = {loss, dL/dexposure_a, dL/dexposure_b, 0}.
// Deterministic: per-block partial sums, the last block adds them in block order.
// ---------------------------------------------------------------------------------------------------------
constexpr int LOSS_THREADS = 256;

struct LossArgs {
    int HW;
    const float *color, *depth, *opacity, *gt_color, *gt_depth, *grad_mask, *exposure;
    float thr, w_rgb, w_depth;
    int flags;
    float *g_color, *g_depth, *g_opacity, *out;
    float *partials;          // [blocks][4]
    unsigned int *ticket;
};

__device__ __forceinline__ float sgnf(float x) { return (float)(x > 0.f) - (float)(x < 0.f); }

__global__ void __launch_bounds__(LOSS_THREADS) fused_loss_kernel(const LossArgs a) {
    __shared__ float s_part[LOSS_THREADS / 32][4];
    __shared__ bool s_last;
    pdl_wait();                                  // launched behind the blend forward (programmatic dependent launch)
    const float ea = a.exposure ? __expf(__ldg(a.exposure)) : 1.f;       // torch.exp in the reference; see tolerance in the tests
    const float eb = a.exposure ? __ldg(a.exposure + 1) : 0.f;
    const float k_rgb = a.w_rgb / (3.f * (float)a.HW), k_d = a.w_depth / (float)a.HW;
    const bool use_opacity = a.flags & LVDGS_LOSS_OPACITY_WEIGHT, opaque_depth = a.flags & LVDGS_LOSS_DEPTH_NEEDS_OPAQUE;
    float loss = 0.f, dea = 0.f, deb = 0.f;
    for (int p = blockIdx.x * LOSS_THREADS + threadIdx.x; p < a.HW; p += gridDim.x * LOSS_THREADS) {
        const float g0 = __ldg(a.gt_color + p), g1 = __ldg(a.gt_color + a.HW + p), g2 = __ldg(a.gt_color + 2 * a.HW + p);
        float m = (g0 + g1 + g2 > a.thr) ? 1.f : 0.f;
        if (a.grad_mask) m *= __ldg(a.grad_mask + p);
        const float op = a.opacity ? __ldg(a.opacity + p) : 1.f;
        const float o = use_opacity ? op : 1.f;
        float abs_sum = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float I = __ldg(a.color + c * a.HW + p);
            const float gt = c == 0 ? g0 : (c == 1 ? g1 : g2);
            const float r = m * fmaf(ea, I, eb) - m * gt;       // the reference's own form: image * mask - gt * mask
            const float s = sgnf(r) * m * o * k_rgb;
            abs_sum += fabsf(r);
            a.g_color[c * a.HW + p] = s * ea;
            dea += s * ea * I;
            deb += s;
        }
        loss += o * abs_sum * k_rgb;
        if (a.g_opacity) a.g_opacity[p] = use_opacity ? abs_sum * k_rgb : 0.f;
        float gd = 0.f;
        if (a.w_depth != 0.f && a.gt_depth) {
            const float gD = __ldg(a.gt_depth + p), D = __ldg(a.depth + p);
            float md = gD > 0.01f ? 1.f : 0.f;
            if (opaque_depth) md *= op > 0.95f ? 1.f : 0.f;
            const float r = D * md - gD * md;
            loss += fabsf(r) * k_d;
            gd = sgnf(r) * md * k_d;
        }
        if (a.g_depth) a.g_depth[p] = gd;
    }
    float v[3] = {loss, dea, deb};
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], d);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { s_part[warp][0] = v[0]; s_part[warp][1] = v[1]; s_part[warp][2] = v[2]; }
    __syncthreads();
    if (threadIdx.x < 3) {
        float t = 0.f;
        for (int w = 0; w < LOSS_THREADS / 32; ++w) t += s_part[w][threadIdx.x];
        a.partials[blockIdx.x * 4 + threadIdx.x] = t;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(a.ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (s_last) {
        // fixed-order sum of the per-block partials by the whole block: thread t adds the blocks b = t/4, t/4 + 64, ...
        // of component t%4, then a fixed tree over the 64 strands -- deterministic, and not 600 dependent loads
        __shared__ float s_fin[LOSS_THREADS];
        const int c = threadIdx.x & 3, j = threadIdx.x >> 2;
        float t = 0.f;
        if (c < 3)
            for (unsigned b = j; b < gridDim.x; b += LOSS_THREADS / 4) t += __ldcg(a.partials + b * 4 + c);
        s_fin[threadIdx.x] = t;
        __syncthreads();
        for (int h = LOSS_THREADS / 8; h >= 1; h >>= 1) {
            if (j < h) s_fin[threadIdx.x] += s_fin[threadIdx.x + 4 * h];
            __syncthreads();
        }
        if (threadIdx.x < 3) a.out[threadIdx.x] = s_fin[threadIdx.x];
        if (threadIdx.x == 3) { a.out[3] = 0.f; *a.ticket = 0u; }
    }
}

constexpr int LOSS_MAX_BLOCKS = 592;     // 4 per SM (296 and 1184 measured the same: the pass is short, its tail is the ticket + final sum)
size_t fused_loss_workspace_bytes() { return align_up(LOSS_MAX_BLOCKS * 4 * sizeof(float)) + 256; }

int launch_fused_loss(int W, int H, const float *color, const float *depth, const float *opacity, const float *gt_color,
                      const float *gt_depth, const float *grad_mask, const float *exposure, float thr, float w_rgb,
                      float w_depth, int flags, float *g_color, float *g_depth, float *g_opacity, float *out, void *ws,
                      size_t ws_bytes, cudaStream_t s) {
    if (ws_bytes < fused_loss_workspace_bytes()) { set_error("fused_loss: workspace too small"); return 1; }
    LossArgs a;
    a.HW = W * H;
    a.color = color; a.depth = depth; a.opacity = opacity; a.gt_color = gt_color; a.gt_depth = gt_depth;
    a.grad_mask = grad_mask; a.exposure = exposure; a.thr = thr; a.w_rgb = w_rgb; a.w_depth = w_depth; a.flags = flags;
    a.g_color = g_color; a.g_depth = g_depth; a.g_opacity = g_opacity; a.out = out;
    a.partials = (float *)ws;
    a.ticket = (unsigned int *)((char *)ws + align_up(LOSS_MAX_BLOCKS * 4 * sizeof(float)));
    const int blocks = min(LOSS_MAX_BLOCKS, ceil_div(a.HW, LOSS_THREADS));
    LVDGS_PRE(s);
    LVDGS_CHECK(launch_after_kernel(fused_loss_kernel, dim3(blocks), dim3(LOSS_THREADS), 0, s, a));
    LVDGS_LAUNCHED(s, "fused_loss");
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// N4: covisibility of two keyframes from their per-Gaussian visibility (utils/slam_frontend.py:1598-1643:
// logical_and / logical_or + count_nonzero, four temporaries and four reductions per pair).  One pass:
// out[4] = {|a|, |b|, |a and b|, |a or b|}; an element counts as visible when non-zero.  elem = bytes per element (1, 4, 8).
// ---------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) covis_kernel(int64_t n, const T *__restrict__ a, const T *__restrict__ b,
                                                    unsigned long long *__restrict__ out) {
    unsigned int ca = 0, cb = 0, ci = 0, cu = 0;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const bool x = a[i] != 0, y = b[i] != 0;
        ca += x; cb += y; ci += x && y; cu += x || y;
    }
    ca = __reduce_add_sync(0xffffffffu, ca); cb = __reduce_add_sync(0xffffffffu, cb);
    ci = __reduce_add_sync(0xffffffffu, ci); cu = __reduce_add_sync(0xffffffffu, cu);
    if ((threadIdx.x & 31) == 0) {
        if (ca) atomicAdd(out + 0, (unsigned long long)ca);
        if (cb) atomicAdd(out + 1, (unsigned long long)cb);
        if (ci) atomicAdd(out + 2, (unsigned long long)ci);
        if (cu) atomicAdd(out + 3, (unsigned long long)cu);
    }
}

int launch_covis(int64_t n, const void *a, const void *b, int elem, unsigned long long *out, cudaStream_t s) {
    LVDGS_CHECK(cudaMemsetAsync(out, 0, 4 * sizeof(unsigned long long), s));
    if (n <= 0) return 0;
    const int blocks = (int)min((int64_t)148 * 8, (n + 255) / 256);
    LVDGS_PRE(s);
    if (elem == 1) covis_kernel<uint8_t><<<blocks, 256, 0, s>>>(n, (const uint8_t *)a, (const uint8_t *)b, out);
    else if (elem == 4) covis_kernel<int32_t><<<blocks, 256, 0, s>>>(n, (const int32_t *)a, (const int32_t *)b, out);
    else if (elem == 8) covis_kernel<long long><<<blocks, 256, 0, s>>>(n, (const long long *)a, (const long long *)b, out);
    else { set_error("covis: element size %d not in {1,4,8}", elem); return 1; }
    LVDGS_LAUNCHED(s, "covis_counts");
    return 0;
}

// n_obs[i] = number of the K masks that see Gaussian i (utils/slam_backend.py:322-325 does this on the CPU after K
// device-to-host copies).  masks: device array of K device pointers.
template <typename T>
__global__ void __launch_bounds__(256) n_obs_kernel(int64_t n, int K, const T *const *__restrict__ masks, int32_t *__restrict__ n_obs) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        int c = 0;
        for (int k = 0; k < K; ++k) c += masks[k][i] != 0;
        n_obs[i] = c;
    }
}

int launch_n_obs(int64_t n, int K, const void *const *masks_dev, int elem, int32_t *n_obs, cudaStream_t s) {
    if (n <= 0) return 0;
    const int blocks = (int)min((int64_t)148 * 8, (n + 255) / 256);
    LVDGS_PRE(s);
    if (elem == 1) n_obs_kernel<uint8_t><<<blocks, 256, 0, s>>>(n, K, (const uint8_t *const *)masks_dev, n_obs);
    else if (elem == 4) n_obs_kernel<int32_t><<<blocks, 256, 0, s>>>(n, K, (const int32_t *const *)masks_dev, n_obs);
    else if (elem == 8) n_obs_kernel<long long><<<blocks, 256, 0, s>>>(n, K, (const long long *const *)masks_dev, n_obs);
    else { set_error("n_obs: element size %d not in {1,4,8}", elem); return 1; }
    LVDGS_LAUNCHED(s, "n_obs");
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// N1: stable row compaction of several row-major float arrays by one keep mask -- GaussianModel.prune_points /
// _prune_optimizer (callers utils/slam_backend.py:128-145, 322-339) index every parameter and both Adam moments with
// the same boolean mask, one torch index kernel + allocation per tensor.  Here: count per 1024-row block, scan the block
// counts, then every block moves its kept rows of ALL arrays (destination rows are contiguous per block, so the
// stores are coalesced).  dst must not alias src.
// ---------------------------------------------------------------------------------------------------------
constexpr int CP_THREADS = 256, CP_ROWS = 1024;
constexpr int CP_MAX_ARRAYS = 16;

__global__ void __launch_bounds__(CP_THREADS) compact_count_kernel(int64_t n, const uint8_t *__restrict__ keep, uint32_t *__restrict__ block_counts) {
    __shared__ uint32_t s_w[CP_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * CP_ROWS;
    uint32_t c = 0;
#pragma unroll
    for (int u = 0; u < CP_ROWS / CP_THREADS; ++u) {
        const int64_t i = base + u * CP_THREADS + threadIdx.x;
        c += i < n && keep[i] != 0;
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < CP_THREADS / 32; ++w) t += s_w[w];
        block_counts[blockIdx.x] = t;
    }
}

// exclusive scan of the block counts in place; total -> counts[nblocks]
__global__ void __launch_bounds__(1024) compact_scan_kernel(int nblocks, uint32_t *__restrict__ counts) {
    __shared__ uint32_t s_ws[32];
    __shared__ uint32_t s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < nblocks; base += 1024) {
        const int b = base + threadIdx.x;
        const uint32_t c = b < nblocks ? counts[b] : 0u;
        uint32_t incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
        if (lane == 31) s_ws[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = s_ws[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, w, d); if (lane >= d) w += t; }
            s_ws[lane] = w;
        }
        __syncthreads();
        const uint32_t excl = s_carry + (warp ? s_ws[warp - 1] : 0u) + incl - c;
        if (b < nblocks) counts[b] = excl;
        __syncthreads();
        if (threadIdx.x == 0) s_carry += s_ws[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) counts[nblocks] = s_carry;
}

struct CompactArrays {
    const float *src[CP_MAX_ARRAYS];
    float *dst[CP_MAX_ARRAYS];
    int width[CP_MAX_ARRAYS];
    int count;
};

__global__ void __launch_bounds__(CP_THREADS) compact_move_kernel(int64_t n, const uint8_t *__restrict__ keep, const uint32_t *__restrict__ block_offsets,
                                                                  const CompactArrays arr) {
    __shared__ uint32_t s_src[CP_ROWS];          // source rows of this block's kept rows, in order
    __shared__ uint32_t s_wbase[CP_THREADS / 32];
    __shared__ uint32_t s_total;
    const int64_t base = (int64_t)blockIdx.x * CP_ROWS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // thread t owns rows base + 4t .. base + 4t + 3 (consecutive, so the kept order is the row order)
    uint32_t flags = 0, c = 0;
#pragma unroll
    for (int u = 0; u < CP_ROWS / CP_THREADS; ++u) {
        const int64_t i = base + (int64_t)threadIdx.x * (CP_ROWS / CP_THREADS) + u;
        if (i < n && keep[i] != 0) { flags |= 1u << u; ++c; }
    }
    uint32_t incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
    if (lane == 31) s_wbase[warp] = incl;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int w = 0; w < CP_THREADS / 32; ++w) { const uint32_t t = s_wbase[w]; s_wbase[w] = run; run += t; }
        s_total = run;
    }
    __syncthreads();
    uint32_t pos = s_wbase[warp] + incl - c;
#pragma unroll
    for (int u = 0; u < CP_ROWS / CP_THREADS; ++u)
        if (flags & (1u << u)) s_src[pos++] = (uint32_t)(threadIdx.x * (CP_ROWS / CP_THREADS) + u);
    __syncthreads();
    const uint32_t total = s_total;
    const size_t dst_row0 = block_offsets[blockIdx.x];
    for (int k = 0; k < arr.count; ++k) {
        const int w = arr.width[k];
        const float *__restrict__ src = arr.src[k] + (size_t)base * w;
        float *__restrict__ dst = arr.dst[k] + dst_row0 * w;
        for (uint32_t e = threadIdx.x; e < total * (uint32_t)w; e += CP_THREADS) {
            const uint32_t r = e / (uint32_t)w, col = e - r * (uint32_t)w;
            dst[e] = src[(size_t)s_src[r] * w + col];
        }
    }
}

size_t compact_workspace_bytes(int64_t n) { return align_up(((size_t)((n + CP_ROWS - 1) / CP_ROWS) + 2) * sizeof(uint32_t)); }

int launch_compact_count(int64_t n, const uint8_t *keep, void *ws, size_t ws_bytes, uint32_t **count_dev, cudaStream_t s) {
    if (ws_bytes < compact_workspace_bytes(n)) { set_error("compact: workspace too small"); return 1; }
    uint32_t *counts = (uint32_t *)ws;
    const int nblocks = (int)((n + CP_ROWS - 1) / CP_ROWS);
    if (nblocks > 0) {
        LVDGS_PRE(s);
        compact_count_kernel<<<nblocks, CP_THREADS, 0, s>>>(n, keep, counts);
        LVDGS_LAUNCHED(s, "compact_count");
    }
    LVDGS_PRE(s);
    compact_scan_kernel<<<1, 1024, 0, s>>>(nblocks, counts);
    LVDGS_LAUNCHED(s, "compact_scan");
    *count_dev = counts + nblocks;
    return 0;
}

int launch_compact_move(int64_t n, const uint8_t *keep, const void *ws, int n_arrays, const float *const *src,
                        float *const *dst, const int32_t *widths, cudaStream_t s) {
    if (n_arrays < 0 || n_arrays > CP_MAX_ARRAYS) { set_error("compact: at most %d arrays per call", CP_MAX_ARRAYS); return 1; }
    const int nblocks = (int)((n + CP_ROWS - 1) / CP_ROWS);
    if (nblocks == 0 || n_arrays == 0) return 0;
    CompactArrays arr;
    arr.count = n_arrays;
    for (int k = 0; k < n_arrays; ++k) {
        if (widths[k] <= 0 || !src[k] || !dst[k] || src[k] == dst[k]) { set_error("compact: bad array %d (in-place is not supported)", k); return 1; }
        arr.src[k] = src[k]; arr.dst[k] = dst[k]; arr.width[k] = widths[k];
    }
    LVDGS_PRE(s);
    compact_move_kernel<<<nblocks, CP_THREADS, 0, s>>>(n, keep, (const uint32_t *)ws, arr);
    LVDGS_LAUNCHED(s, "compact_move");
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// Rows a15 / a16 on the device: one tracking iteration's tail.  The reference runs torch.optim.Adam on
// (cam_rot_delta, cam_trans_delta, exposure_a, exposure_b) and then update_pose (utils/slam_frontend.py:1466-1521,
// utils/pose_utils.py:56-87): ~60 tiny torch kernels and two host synchronisations (`if angle < 1e-5`, `if converged`)
// per iteration.  Here: one single-thread kernel on the camera's device-resident state block (lvdgs_pose_state):
//   Adam step (the deltas are zero before every step, as update_pose resets them) -> tau = [trans_delta; rot_delta]
//   -> T_w2c <- SE3_exp(tau) T_w2c -> world_view_transform, full_proj_transform, camera_center refreshed in the layout
//   the rasterizer reads (utils/camera_utils.py:106-120) -> converged = |tau| < threshold.
// ---------------------------------------------------------------------------------------------------------
__global__ void pose_step_kernel(lvdgs_pose_state *st, const float *__restrict__ g_tau, const float *__restrict__ g_exposure,
                                 float lr_rot, float lr_trans, float lr_exp, float beta1, float beta2, float eps,
                                 float bc1, float bc2_sqrt, float threshold) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    pdl_wait();                                  // launched behind the preprocess backward (programmatic dependent launch)
    // gradients in parameter order: rot_delta (theta), trans_delta (rho), exposure a, b
    float g[8];
#pragma unroll
    for (int k = 0; k < 3; ++k) { g[k] = g_tau[3 + k]; g[3 + k] = g_tau[k]; }
    g[6] = g_exposure ? g_exposure[0] : 0.f;
    g[7] = g_exposure ? g_exposure[1] : 0.f;
    float delta[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float lr = k < 3 ? lr_rot : (k < 6 ? lr_trans : lr_exp);
        const float m = beta1 * st->adam_m[k] + (1.f - beta1) * g[k];
        const float v = beta2 * st->adam_v[k] + (1.f - beta2) * g[k] * g[k];
        st->adam_m[k] = m; st->adam_v[k] = v;
        const float denom = sqrtf(v) / bc2_sqrt + eps;                 // torch.optim.Adam's order of operations
        delta[k] = -(lr / bc1) * (m / denom);
    }
    if (g_exposure) { st->exposure[0] += delta[6]; st->exposure[1] += delta[7]; }
    const float th[3] = {delta[0], delta[1], delta[2]}, rho[3] = {delta[3], delta[4], delta[5]};
    // SE3_exp(tau), utils/pose_utils.py:22-68
    const float W[9] = {0.f, -th[2], th[1], th[2], 0.f, -th[0], -th[1], th[0], 0.f};
    float W2[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) W2[3 * i + j] = W[3 * i] * W[j] + W[3 * i + 1] * W[3 + j] + W[3 * i + 2] * W[6 + j];
    const float angle = sqrtf(th[0] * th[0] + th[1] * th[1] + th[2] * th[2]);
    float a1, a2, b1, b2;            // R = I + a1 W + a2 W2,  V = I + b1 W + b2 W2
    if (angle < 1e-5f) { a1 = 1.f; a2 = 0.5f; b1 = 0.5f; b2 = 1.f / 6.f; }
    else {
        const float s = sinf(angle), c = cosf(angle);
        a1 = s / angle; a2 = (1.f - c) / (angle * angle);
        b1 = a2; b2 = (angle - s) / (angle * angle * angle);
    }
    float Rd[9], Vd[9], t[3];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const float I = (k % 4 == 0) ? 1.f : 0.f;
        Rd[k] = I + a1 * W[k] + a2 * W2[k];
        Vd[k] = I + b1 * W[k] + b2 * W2[k];
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) t[i] = Vd[3 * i] * rho[0] + Vd[3 * i + 1] * rho[1] + Vd[3 * i + 2] * rho[2];
    // new_w2c = SE3_exp(tau) @ T_w2c
    float Rn[9], Tn[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) Rn[3 * i + j] = Rd[3 * i] * st->R[j] + Rd[3 * i + 1] * st->R[3 + j] + Rd[3 * i + 2] * st->R[6 + j];
        Tn[i] = Rd[3 * i] * st->T[0] + Rd[3 * i + 1] * st->T[1] + Rd[3 * i + 2] * st->T[2] + t[i];
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) st->R[k] = Rn[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) st->T[k] = Tn[k];
    // world_view_transform = [R T; 0 1]^T (flat row-major), full_proj = world_view_transform @ projection_matrix
    float V4[16];
#pragma unroll
    for (int r = 0; r < 3; ++r) { V4[4 * r] = Rn[r]; V4[4 * r + 1] = Rn[3 + r]; V4[4 * r + 2] = Rn[6 + r]; V4[4 * r + 3] = 0.f; }
    V4[12] = Tn[0]; V4[13] = Tn[1]; V4[14] = Tn[2]; V4[15] = 1.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) st->view[k] = V4[k];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) acc += V4[4 * i + k] * st->proj_raw[4 * k + j];
            st->proj[4 * i + j] = acc;
        }
    // camera centre = -R^T T
#pragma unroll
    for (int j = 0; j < 3; ++j) st->campos[j] = -(Rn[j] * Tn[0] + Rn[3 + j] * Tn[1] + Rn[6 + j] * Tn[2]);
    const float tn = sqrtf(th[0] * th[0] + th[1] * th[1] + th[2] * th[2] + rho[0] * rho[0] + rho[1] * rho[1] + rho[2] * rho[2]);
    st->tau_norm = tn;
    st->converged = tn < threshold ? 1 : 0;
    st->step += 1;
}

int launch_pose_step(lvdgs_pose_state *state, const float *g_tau, const float *g_exposure, float lr_rot, float lr_trans,
                     float lr_exp, double beta1, double beta2, double eps, int step, float threshold, cudaStream_t s) {
    const float bc1 = (float)(1.0 - pow(beta1, (double)step));
    const float bc2_sqrt = (float)sqrt(1.0 - pow(beta2, (double)step));
    LVDGS_PRE(s);
    LVDGS_CHECK(launch_after_kernel(pose_step_kernel, dim3(1), dim3(32), 0, s, state, g_tau, g_exposure, lr_rot, lr_trans, lr_exp, (float)beta1,
                                    (float)beta2, (float)eps, bc1, bc2_sqrt, threshold));
    LVDGS_LAUNCHED(s, "pose_step");
    return 0;
}

}  // namespace lvdgs

namespace lvdgs {

// ---------------------------------------------------------------------------------------------------------
// N1, densification half: rows idx[j] of several row-major float arrays -> row j of the destinations (dst may be the tail
// of the same allocation as src: densify_and_clone / densify_and_split append copies of the selected Gaussians to every
// parameter tensor, utils/slam_backend.py:359-376 via GaussianModel.densify_and_prune).  One launch for all arrays;
// consecutive threads move consecutive floats of a destination row (coalesced stores, row-granular gathers).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gather_rows_kernel(int64_t n_idx, const int64_t *__restrict__ idx, int64_t n_src_rows, const CompactArrays arr) {
    for (int k = 0; k < arr.count; ++k) {
        const int w = arr.width[k];
        const int64_t total = n_idx * w;
        for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (int64_t)gridDim.x * 256) {
            const int64_t r = e / w;
            const int col = (int)(e - r * w);
            const int64_t s = idx[r];
            if (s >= 0 && s < n_src_rows) arr.dst[k][e] = arr.src[k][s * w + col];
        }
    }
}

int launch_gather_rows(int64_t n_idx, const int64_t *idx, int64_t n_src_rows, int n_arrays, const float *const *src,
                       float *const *dst, const int32_t *widths, cudaStream_t s) {
    if (n_arrays < 0 || n_arrays > CP_MAX_ARRAYS) { set_error("gather_rows: at most %d arrays per call", CP_MAX_ARRAYS); return 1; }
    if (n_idx <= 0 || n_arrays == 0) return 0;
    CompactArrays arr;
    arr.count = n_arrays;
    int wmax = 1;
    for (int k = 0; k < n_arrays; ++k) {
        if (widths[k] <= 0 || !src[k] || !dst[k]) { set_error("gather_rows: bad array %d", k); return 1; }
        arr.src[k] = src[k]; arr.dst[k] = dst[k]; arr.width[k] = widths[k];
        wmax = max(wmax, widths[k]);
    }
    const int blocks = (int)min((int64_t)148 * 8, (n_idx * wmax + 255) / 256);
    LVDGS_PRE(s);
    gather_rows_kernel<<<blocks, 256, 0, s>>>(n_idx, idx, n_src_rows, arr);
    LVDGS_LAUNCHED(s, "gather_rows");
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// Row N1: GaussianModel's parametrisation on the device.  The reference optimises RAW parameters -- logit opacity,
// log scale, un-normalised quaternion -- and hands the rasterizer their activations (get_opacity = sigmoid,
// get_scaling = exp, get_rotation = normalize; SURVEY.md A.0, callers utils/slam_backend.py:98,184,277).
// activate: raw -> rasterizer inputs, one thread per Gaussian.  activation_backward: gradients with respect to the
// activated values (what lvdgs_rasterize_backward produces) -> gradients with respect to the raw parameters, in place:
//   d/d raw_o = g o (1 - o),   d/d raw_s = g s,   d/d raw_q = (g - q^ <q^, g>) / |raw_q|.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gaussian_activate_kernel(int64_t P, const float *__restrict__ raw_o, const float *__restrict__ raw_s,
                                                                const float4 *__restrict__ raw_q, float *__restrict__ o, float *__restrict__ sc,
                                                                float4 *__restrict__ q) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= P) return;
    o[i] = 1.f / (1.f + expf(-raw_o[i]));
#pragma unroll
    for (int k = 0; k < 3; ++k) sc[3 * i + k] = expf(raw_s[3 * i + k]);
    const float4 r = raw_q[i];
    const float n = fmaxf(sqrtf(r.x * r.x + r.y * r.y + r.z * r.z + r.w * r.w), 1e-12f);     // torch.nn.functional.normalize eps
    q[i] = make_float4(r.x / n, r.y / n, r.z / n, r.w / n);
}

__global__ void __launch_bounds__(256) gaussian_activation_backward_kernel(int64_t P, const float *__restrict__ o, const float *__restrict__ sc,
                                                                           const float4 *__restrict__ q, const float4 *__restrict__ raw_q,
                                                                           float *__restrict__ g_o, float *__restrict__ g_s, float4 *__restrict__ g_q) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= P) return;
    const float oi = o[i];
    g_o[i] *= oi * (1.f - oi);
#pragma unroll
    for (int k = 0; k < 3; ++k) g_s[3 * i + k] *= sc[3 * i + k];
    const float4 qi = q[i], r = raw_q[i], g = g_q[i];
    const float inv = 1.f / fmaxf(sqrtf(r.x * r.x + r.y * r.y + r.z * r.z + r.w * r.w), 1e-12f);
    const float d = qi.x * g.x + qi.y * g.y + qi.z * g.z + qi.w * g.w;
    g_q[i] = make_float4((g.x - qi.x * d) * inv, (g.y - qi.y * d) * inv, (g.z - qi.z * d) * inv, (g.w - qi.w * d) * inv);
}

int launch_gaussian_activate(int64_t P, const float *raw_o, const float *raw_s, const float *raw_q, float *o, float *sc, float *q,
                             cudaStream_t s) {
    if (P <= 0) return 0;
    LVDGS_PRE(s);
    gaussian_activate_kernel<<<(unsigned)((P + 255) / 256), 256, 0, s>>>(P, raw_o, raw_s, reinterpret_cast<const float4 *>(raw_q), o, sc,
                                                                          reinterpret_cast<float4 *>(q));
    LVDGS_LAUNCHED(s, "gaussian_activate");
    return 0;
}

int launch_gaussian_activation_backward(int64_t P, const float *o, const float *sc, const float *q, const float *raw_q, float *g_o,
                                        float *g_s, float *g_q, cudaStream_t s) {
    if (P <= 0) return 0;
    LVDGS_PRE(s);
    gaussian_activation_backward_kernel<<<(unsigned)((P + 255) / 256), 256, 0, s>>>(P, o, sc, reinterpret_cast<const float4 *>(q),
                                                                                     reinterpret_cast<const float4 *>(raw_q), g_o, g_s,
                                                                                     reinterpret_cast<float4 *>(g_q));
    LVDGS_LAUNCHED(s, "gaussian_activation_backward");
    return 0;
}

}  // namespace lvdgs
