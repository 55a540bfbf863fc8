// ssim_loss.cu -- row N3, the loss LVD-GS's mapping actually runs when a keyframe carries a static mask
// (utils/slam_backend.py:199-261): pixels of dynamic objects are painted with the background colour in both the render
// and the ground truth, then
//     loss = (1 - lambda) * mean |mi - mg|  +  lambda * (1 - mean SSIM(mi, mg))  +  depth_lambda * mean_{dm} |D - mono|,
//     dm = static_mask & (mono > 0) & (D > 0),
// with gaussian_splatting.utils.loss_utils.l1_loss / ssim (11x11 Gaussian window, sigma 1.5, zero-padded depthwise
// convolutions, C1 = 0.01^2, C2 = 0.03^2).  In the reference that is ~40 torch kernels forward (five 3-channel conv2d, the
// mask fills, boolean gathers) and as many backward, with a dozen full-image temporaries.  Here: two tiled kernels.
//   ssim_forward : per 16x16 tile and channel, the masked images with an 5-pixel halo in shared memory, separable 11-tap
//                  filtering of (x, y, x^2, y^2, x y), the SSIM value and the three derivative maps
//                  A = dS/dmu1 (total), B = dS/dE[x^2], C = dS/dE[xy]; per-block partial sums of SSIM, |x - y|, the depth
//                  term and its pixel count (deterministic: fixed-order final sum by the last block).
//   ssim_backward: dL/dx = kS * (w * A + 2 x (w * B) + y (w * C)) + kL * sign(x - y)   (w * . = the same filter), zero on
//                  masked-out pixels; dL/dD = depth_lambda * sign(D - mono) / count on dm.
#include "common.cuh"

namespace lvdgs {

constexpr int SS_T = 16, SS_R = 5, SS_W = SS_T + 2 * SS_R;      // tile, window radius, tile + halo
constexpr int SS_THREADS = SS_T * SS_T;

struct SsimWindow { float g[2 * SS_R + 1]; };

// exp(-(k-5)^2 / (2 * 1.5^2)) normalised, accumulated in float like torch.Tensor([...]) / sum()
static SsimWindow make_window() {
    SsimWindow w;
    float s = 0.f;
    for (int k = 0; k <= 2 * SS_R; ++k) { w.g[k] = (float)exp(-(double)((k - SS_R) * (k - SS_R)) / (2.0 * 1.5 * 1.5)); s += w.g[k]; }
    for (int k = 0; k <= 2 * SS_R; ++k) w.g[k] /= s;
    return w;
}

struct SsimArgs {
    int W, H;
    const float *image, *gt, *bg, *depth, *mono;
    const uint8_t *mask;                 // [H*W] non-zero = static (kept) pixel; NULL = everything kept
    float lambda_dssim, depth_lambda;
    float *maps;                         // [3 maps][3 channels][H*W]
    float *partials;                     // [blocks][4]: ssim sum, |x - y| sum, depth |.| sum, depth count
    unsigned int *ticket;
    float *out;                          // [8]: loss, ssim mean, l1 mean, depth mean, depth count, -, -, -
    float *g_image, *g_depth;
    SsimWindow win;
};

__device__ __forceinline__ float sgn1(float x) { return (float)(x > 0.f) - (float)(x < 0.f); }

__global__ void __launch_bounds__(SS_THREADS) ssim_forward_kernel(const SsimArgs a) {
    __shared__ float s_x[SS_W][SS_W + 1], s_y[SS_W][SS_W + 1];
    __shared__ float s_h[5][SS_W][SS_T + 1];
    __shared__ float s_part[SS_THREADS / 32][4];
    __shared__ bool s_last;
    const int c = blockIdx.z, x0 = blockIdx.x * SS_T, y0 = blockIdx.y * SS_T;
    const int tx = threadIdx.x % SS_T, ty = threadIdx.x / SS_T;
    const size_t HW = (size_t)a.H * a.W;
    const float bgc = __ldg(a.bg + c);
    for (int i = threadIdx.x; i < SS_W * SS_W; i += SS_THREADS) {
        const int r = i / SS_W, q = i % SS_W, gy = y0 + r - SS_R, gx = x0 + q - SS_R;
        float vx = 0.f, vy = 0.f;                                     // zero padding of conv2d outside the image
        if (gx >= 0 && gx < a.W && gy >= 0 && gy < a.H) {
            const size_t p = (size_t)gy * a.W + gx;
            const bool keep = !a.mask || a.mask[p];
            vx = keep ? __ldg(a.image + c * HW + p) : bgc;
            vy = keep ? __ldg(a.gt + c * HW + p) : bgc;
        }
        s_x[r][q] = vx; s_y[r][q] = vy;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < SS_W * SS_T; i += SS_THREADS) {        // horizontal pass
        const int r = i / SS_T, q = i % SS_T;
        float m1 = 0.f, m2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
        for (int k = 0; k <= 2 * SS_R; ++k) {
            const float w = a.win.g[k], vx = s_x[r][q + k], vy = s_y[r][q + k];
            m1 += w * vx; m2 += w * vy; e11 += w * vx * vx; e22 += w * vy * vy; e12 += w * vx * vy;
        }
        s_h[0][r][q] = m1; s_h[1][r][q] = m2; s_h[2][r][q] = e11; s_h[3][r][q] = e22; s_h[4][r][q] = e12;
    }
    __syncthreads();
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    const int gx = x0 + tx, gy = y0 + ty;
    if (gx < a.W && gy < a.H) {
        float m1 = 0.f, m2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
        for (int k = 0; k <= 2 * SS_R; ++k) {                             // vertical pass
            const float w = a.win.g[k];
            m1 += w * s_h[0][ty + k][tx]; m2 += w * s_h[1][ty + k][tx]; e11 += w * s_h[2][ty + k][tx];
            e22 += w * s_h[3][ty + k][tx]; e12 += w * s_h[4][ty + k][tx];
        }
        const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
        const float a1 = 2.f * m1 * m2 + C1, a2 = 2.f * (e12 - m1 * m2) + C2;
        const float b1 = m1 * m1 + m2 * m2 + C1, b2 = (e11 - m1 * m1) + (e22 - m2 * m2) + C2;
        const float inv = 1.f / (b1 * b2);
        const float ssim = a1 * a2 * inv;
        const size_t p = (size_t)gy * a.W + gx, o = c * HW + p;
        a.maps[o] = 2.f * m2 * (a2 - a1) * inv - 2.f * m1 * ssim * (b2 - b1) * inv;      // A: dS/dmu1, total
        a.maps[3 * HW + o] = -ssim / b2;                                                   // B: dS/dE[x^2]
        a.maps[6 * HW + o] = 2.f * a1 * inv;                                               // C: dS/dE[xy]
        v[0] = ssim;
        v[1] = fabsf(s_x[ty + SS_R][tx + SS_R] - s_y[ty + SS_R][tx + SS_R]);
        if (c == 0 && a.depth && a.mono) {
            const float D = __ldg(a.depth + p), mo = __ldg(a.mono + p);
            if ((!a.mask || a.mask[p]) && mo > 0.f && D > 0.f) { v[2] = fabsf(D - mo); v[3] = 1.f; }
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], d);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { s_part[warp][0] = v[0]; s_part[warp][1] = v[1]; s_part[warp][2] = v[2]; s_part[warp][3] = v[3]; }
    __syncthreads();
    const unsigned nblocks = gridDim.x * gridDim.y * gridDim.z;
    const unsigned bid = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    if (threadIdx.x < 4) {
        float t = 0.f;
        for (int w = 0; w < SS_THREADS / 32; ++w) t += s_part[w][threadIdx.x];
        a.partials[(size_t)bid * 4 + threadIdx.x] = t;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(a.ticket, 1u) == nblocks - 1;
    __syncthreads();
    if (s_last) {       // fixed-order sum of the per-block partials: deterministic
        __shared__ double s_fin[SS_THREADS];
        const int k = threadIdx.x & 3, j = threadIdx.x >> 2;
        double t = 0.0;
        for (unsigned b = j; b < nblocks; b += SS_THREADS / 4) t += (double)__ldcg(a.partials + (size_t)b * 4 + k);
        s_fin[threadIdx.x] = t;
        __syncthreads();
        for (int h = SS_THREADS / 8; h >= 1; h >>= 1) {
            if (j < h) s_fin[threadIdx.x] += s_fin[threadIdx.x + 4 * h];
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            const double n = 3.0 * (double)HW;
            const double ssim_mean = s_fin[0] / n, l1_mean = s_fin[1] / n, cnt = s_fin[3];
            const double d_mean = cnt > 0.0 ? s_fin[2] / cnt : 0.0;
            a.out[0] = (float)((1.0 - a.lambda_dssim) * l1_mean + a.lambda_dssim * (1.0 - ssim_mean) + (cnt > 0.0 ? a.depth_lambda * d_mean : 0.0));
            a.out[1] = (float)ssim_mean; a.out[2] = (float)l1_mean; a.out[3] = (float)d_mean; a.out[4] = (float)cnt;
            *a.ticket = 0u;
        }
    }
}

__global__ void __launch_bounds__(SS_THREADS) ssim_backward_kernel(const SsimArgs a) {
    __shared__ float s_m[3][SS_W][SS_W + 1];
    __shared__ float s_h[3][SS_W][SS_T + 1];
    const int c = blockIdx.z, x0 = blockIdx.x * SS_T, y0 = blockIdx.y * SS_T;
    const int tx = threadIdx.x % SS_T, ty = threadIdx.x / SS_T;
    const size_t HW = (size_t)a.H * a.W;
    for (int i = threadIdx.x; i < SS_W * SS_W; i += SS_THREADS) {
        const int r = i / SS_W, q = i % SS_W, gy = y0 + r - SS_R, gx = x0 + q - SS_R;
        const bool in = gx >= 0 && gx < a.W && gy >= 0 && gy < a.H;
        const size_t o = c * HW + (size_t)(in ? gy : 0) * a.W + (in ? gx : 0);
#pragma unroll
        for (int k = 0; k < 3; ++k) s_m[k][r][q] = in ? a.maps[k * 3 * HW + o] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < SS_W * SS_T; i += SS_THREADS) {
        const int r = i / SS_T, q = i % SS_T;
        float t0 = 0.f, t1 = 0.f, t2 = 0.f;
#pragma unroll
        for (int k = 0; k <= 2 * SS_R; ++k) { const float w = a.win.g[k]; t0 += w * s_m[0][r][q + k]; t1 += w * s_m[1][r][q + k]; t2 += w * s_m[2][r][q + k]; }
        s_h[0][r][q] = t0; s_h[1][r][q] = t1; s_h[2][r][q] = t2;
    }
    __syncthreads();
    const int gx = x0 + tx, gy = y0 + ty;
    if (gx >= a.W || gy >= a.H) return;
    float cA = 0.f, cB = 0.f, cC = 0.f;
#pragma unroll
    for (int k = 0; k <= 2 * SS_R; ++k) { const float w = a.win.g[k]; cA += w * s_h[0][ty + k][tx]; cB += w * s_h[1][ty + k][tx]; cC += w * s_h[2][ty + k][tx]; }
    const size_t p = (size_t)gy * a.W + gx;
    const bool keep = !a.mask || a.mask[p];
    const float n = 3.f * (float)HW;
    float g = 0.f;
    if (keep) {       // a masked-out pixel is a constant (the background colour): no gradient reaches the render there
        const float x = __ldg(a.image + c * HW + p), y = __ldg(a.gt + c * HW + p);
        g = -a.lambda_dssim / n * (cA + 2.f * x * cB + y * cC) + (1.f - a.lambda_dssim) / n * sgn1(x - y);
    }
    a.g_image[c * HW + p] = g;
    if (c == 0 && a.g_depth) {
        float gd = 0.f;
        if (a.depth && a.mono) {
            const float D = __ldg(a.depth + p), mo = __ldg(a.mono + p), cnt = a.out[4];
            if (keep && mo > 0.f && D > 0.f && cnt > 0.f) gd = a.depth_lambda * sgn1(D - mo) / cnt;
        }
        a.g_depth[p] = gd;
    }
}

size_t masked_ssim_workspace_bytes(int W, int H) {
    const size_t HW = (size_t)W * H;
    const size_t blocks = (size_t)ceil_div(W, SS_T) * ceil_div(H, SS_T) * 3;
    return align_up(9 * HW * sizeof(float)) + align_up(blocks * 4 * sizeof(float)) + 256;
}

int launch_masked_ssim_loss(int W, int H, const float *image, const float *gt, const uint8_t *mask, const float *bg,
                            const float *depth, const float *mono, float lambda_dssim, float depth_lambda, float *g_image,
                            float *g_depth, float *out, void *ws, size_t ws_bytes, cudaStream_t s) {
    if (ws_bytes < masked_ssim_workspace_bytes(W, H)) { set_error("masked_ssim_loss: workspace too small"); return 1; }
    static const SsimWindow win = make_window();
    const size_t HW = (size_t)W * H;
    const dim3 grid(ceil_div(W, SS_T), ceil_div(H, SS_T), 3);
    SsimArgs a;
    a.W = W; a.H = H; a.image = image; a.gt = gt; a.bg = bg; a.depth = depth; a.mono = mono; a.mask = mask;
    a.lambda_dssim = lambda_dssim; a.depth_lambda = depth_lambda;
    a.maps = (float *)ws;
    a.partials = (float *)((char *)ws + align_up(9 * HW * sizeof(float)));
    a.ticket = (unsigned int *)((char *)a.partials + align_up((size_t)grid.x * grid.y * grid.z * 4 * sizeof(float)));
    a.out = out; a.g_image = g_image; a.g_depth = g_depth; a.win = win;
    LVDGS_PRE(s);
    ssim_forward_kernel<<<grid, SS_THREADS, 0, s>>>(a);
    LVDGS_LAUNCHED(s, "ssim_forward");
    LVDGS_PRE(s);
    ssim_backward_kernel<<<grid, SS_THREADS, 0, s>>>(a);
    LVDGS_LAUNCHED(s, "ssim_backward");
    return 0;
}

}  // namespace lvdgs
