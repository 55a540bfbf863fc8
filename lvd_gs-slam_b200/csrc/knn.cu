// knn.cu -- K12: simple_knn.distCUDA2 replacement (reached from
// GaussianModel.extend_from_pcd_seq via utils/slam_backend.py:75-78): for every point the mean of the squared
// distances to its three nearest neighbours, EXACT.
//
// Same published algorithm as the reference's plugin, rebuilt on this library's own pieces:
//   1. bounding box of the cloud (block reduction + ordered-int atomics);
//   2. 30-bit Morton code per point (10 bits per axis inside the box);
//   3. stable sort of (code, index) with the onesweep radix sort of radix_sort.cu (30 key bits = 4 passes);
//   4. the points are gathered into Morton order as float4 (so that a box is one contiguous, coalesced range) and
//      every run of 256 consecutive points gets its axis-aligned bounding box;
//   5. per point (one thread each, neighbours in Morton order share a warp and therefore visit the same boxes):
//      seed the three best squared distances from the +-3 neighbours in Morton order, then scan every box whose
//      AABB is not farther than the current third-best distance.  The pruning is conservative, so the result is the
//      exact 3-NN mean; only float rounding of the squared distances differs from a brute-force evaluation.
// Small clouds (P <= 2048) skip the machinery and use the all-pairs kernel.
#include "common.cuh"
#include <float.h>

namespace lvdgs {

constexpr int KNN_THREADS = 256;
constexpr int KNN_BOX = 256;

// ---------------- small clouds: all pairs, tiled through shared memory ----------------
__global__ void __launch_bounds__(KNN_THREADS) dist2_bruteforce_kernel(int P, const float *__restrict__ pts,
                                                                       float *__restrict__ out) {
    __shared__ float4 s_p[KNN_THREADS];
    const int i = blockIdx.x * KNN_THREADS + threadIdx.x;
    float3 q = make_float3(0.f, 0.f, 0.f);
    if (i < P) q = make_float3(pts[3 * (size_t)i], pts[3 * (size_t)i + 1], pts[3 * (size_t)i + 2]);
    float b0 = FLT_MAX, b1 = FLT_MAX, b2 = FLT_MAX;
    for (int base = 0; base < P; base += KNN_THREADS) {
        const int j = base + threadIdx.x;
        __syncthreads();
        s_p[threadIdx.x] = j < P ? make_float4(pts[3 * (size_t)j], pts[3 * (size_t)j + 1], pts[3 * (size_t)j + 2], 0.f)
                                 : make_float4(FLT_MAX, FLT_MAX, FLT_MAX, 0.f);
        __syncthreads();
        const int nb = min(KNN_THREADS, P - base);
#pragma unroll 8
        for (int k = 0; k < nb; ++k) {
            const float4 c = s_p[k];
            const float dx = c.x - q.x, dy = c.y - q.y, dz = c.z - q.z;
            const float d = dx * dx + dy * dy + dz * dz;
            if (d < b2 && base + k != i) {
                if (d < b1) {
                    b2 = b1;
                    if (d < b0) { b1 = b0; b0 = d; } else b1 = d;
                } else b2 = d;
            }
        }
    }
    if (i < P) out[i] = (b0 + b1 + b2) / 3.f;
}

// ---------------- large clouds ----------------
// order-preserving float <-> uint map so that atomicMin / atomicMax work on floats of either sign
__device__ __forceinline__ uint32_t f2ord(float f) { const uint32_t u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float ord2f(uint32_t u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

__global__ void __launch_bounds__(KNN_THREADS) knn_bbox_kernel(int P, const float *__restrict__ pts, uint32_t *__restrict__ bbox) {
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = blockIdx.x * KNN_THREADS + threadIdx.x; i < P; i += gridDim.x * KNN_THREADS)
#pragma unroll
        for (int a = 0; a < 3; ++a) { const float v = pts[3 * (size_t)i + a]; mn[a] = fminf(mn[a], v); mx[a] = fmaxf(mx[a], v); }
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], d));
            mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], d));
        }
    if ((threadIdx.x & 31) == 0)
#pragma unroll
        for (int a = 0; a < 3; ++a) { atomicMin(bbox + a, f2ord(mn[a])); atomicMax(bbox + 3 + a, f2ord(mx[a])); }
}

__device__ __forceinline__ uint32_t spread10(uint32_t x) {   // 10 bits -> every third bit
    x = (x | (x << 16)) & 0x030000FFu;
    x = (x | (x << 8)) & 0x0300F00Fu;
    x = (x | (x << 4)) & 0x030C30C3u;
    x = (x | (x << 2)) & 0x09249249u;
    return x;
}

__global__ void __launch_bounds__(KNN_THREADS) knn_morton_kernel(int P, const float *__restrict__ pts, const uint32_t *__restrict__ bbox,
                                                                 uint64_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    const int i = blockIdx.x * KNN_THREADS + threadIdx.x;
    if (i >= P) return;
    uint32_t code = 0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float lo = ord2f(bbox[a]), hi = ord2f(bbox[3 + a]);
        const float ext = hi - lo;
        const float t = ext > 0.f ? (pts[3 * (size_t)i + a] - lo) / ext : 0.f;
        const uint32_t q = (uint32_t)fminf(fmaxf(t * 1023.f, 0.f), 1023.f);
        code |= spread10(q) << a;
    }
    keys[i] = code;
    vals[i] = (uint32_t)i;
}

// gather into Morton order + per-box AABB (one CTA per box)
__global__ void __launch_bounds__(KNN_BOX) knn_boxes_kernel(int P, const float *__restrict__ pts, const uint32_t *__restrict__ order,
                                                            float4 *__restrict__ sorted, float *__restrict__ boxes) {
    __shared__ float s_mn[32][3], s_mx[32][3];
    const int i = blockIdx.x * KNN_BOX + threadIdx.x;
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    if (i < P) {
        const uint32_t o = order[i];
        const float x = pts[3 * (size_t)o], y = pts[3 * (size_t)o + 1], z = pts[3 * (size_t)o + 2];
        sorted[i] = make_float4(x, y, z, __uint_as_float(o));
        mn[0] = mx[0] = x; mn[1] = mx[1] = y; mn[2] = mx[2] = z;
    }
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], d));
            mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], d));
        }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0)
#pragma unroll
        for (int a = 0; a < 3; ++a) { s_mn[warp][a] = mn[a]; s_mx[warp][a] = mx[a]; }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            float lo = lane < KNN_BOX / 32 ? s_mn[lane][a] : FLT_MAX, hi = lane < KNN_BOX / 32 ? s_mx[lane][a] : -FLT_MAX;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, d));
                hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, d));
            }
            if (lane == 0) { boxes[6 * blockIdx.x + a] = lo; boxes[6 * blockIdx.x + 3 + a] = hi; }
        }
    }
}

__device__ __forceinline__ void best3_insert(float d, float &b0, float &b1, float &b2) {
    if (d < b2) {
        if (d < b1) {
            b2 = b1;
            if (d < b0) { b1 = b0; b0 = d; } else b1 = d;
        } else b2 = d;
    }
}

// second level: AABB of every run of KNN_SUPER consecutive boxes, so that a query rejects 32 boxes with one test
constexpr int KNN_SUPER = 32;
__global__ void __launch_bounds__(KNN_THREADS) knn_super_kernel(int nboxes, const float *__restrict__ boxes, float *__restrict__ supers) {
    const int sidx = blockIdx.x * KNN_THREADS + threadIdx.x;
    if (sidx * KNN_SUPER >= nboxes) return;
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int b = sidx * KNN_SUPER; b < min(nboxes, (sidx + 1) * KNN_SUPER); ++b)
#pragma unroll
        for (int a = 0; a < 3; ++a) { mn[a] = fminf(mn[a], boxes[6 * b + a]); mx[a] = fmaxf(mx[a], boxes[6 * b + 3 + a]); }
#pragma unroll
    for (int a = 0; a < 3; ++a) { supers[6 * sidx + a] = mn[a]; supers[6 * sidx + 3 + a] = mx[a]; }
}

__device__ __forceinline__ float aabb_dist2(const float *__restrict__ bx, float4 q) {
    const float ex = fmaxf(fmaxf(__ldg(bx + 0) - q.x, q.x - __ldg(bx + 3)), 0.f);
    const float ey = fmaxf(fmaxf(__ldg(bx + 1) - q.y, q.y - __ldg(bx + 4)), 0.f);
    const float ez = fmaxf(fmaxf(__ldg(bx + 2) - q.z, q.z - __ldg(bx + 5)), 0.f);
    return ex * ex + ey * ey + ez * ez;
}

__global__ void __launch_bounds__(KNN_THREADS) knn_search_kernel(int P, int nboxes, const float4 *__restrict__ sorted,
                                                                 const float *__restrict__ boxes, const float *__restrict__ supers,
                                                                 float *__restrict__ out) {
    const int i = blockIdx.x * KNN_THREADS + threadIdx.x;
    if (i >= P) return;
    const float4 q = sorted[i];
    float b0 = FLT_MAX, b1 = FLT_MAX, b2 = FLT_MAX;
    for (int j = max(0, i - 3); j <= min(P - 1, i + 3); ++j) {
        if (j == i) continue;
        const float4 c = sorted[j];
        const float dx = c.x - q.x, dy = c.y - q.y, dz = c.z - q.z;
        best3_insert(dx * dx + dy * dy + dz * dz, b0, b1, b2);
    }
    // the seed only provides a rejection bound (the true third-nearest distance cannot exceed it); the best three are
    // then collected from scratch out of the boxes, so no neighbour is counted twice
    const float bound = b2;
    b0 = b1 = b2 = FLT_MAX;
    const int nsuper = (nboxes + KNN_SUPER - 1) / KNN_SUPER;
    const int own = i / KNN_BOX;
    {   // the query's own box first: it almost always holds the true neighbours, so the bound is tight for all the others
        const int j0 = own * KNN_BOX, j1 = min(P, j0 + KNN_BOX);
        for (int j = j0; j < j1; ++j) {
            if (j == i) continue;
            const float4 c = __ldg(sorted + j);
            const float dx = c.x - q.x, dy = c.y - q.y, dz = c.z - q.z;
            best3_insert(dx * dx + dy * dy + dz * dz, b0, b1, b2);
        }
    }
    for (int sb = 0; sb < nsuper; ++sb) {
        if (aabb_dist2(supers + 6 * sb, q) > fminf(bound, b2)) continue;
        for (int b = sb * KNN_SUPER; b < min(nboxes, (sb + 1) * KNN_SUPER); ++b) {
            if (b == own || aabb_dist2(boxes + 6 * b, q) > fminf(bound, b2)) continue;
            const int j0 = b * KNN_BOX, j1 = min(P, j0 + KNN_BOX);
            for (int j = j0; j < j1; ++j) {
                if (j == i) continue;
                const float4 c = __ldg(sorted + j);
                const float dx = c.x - q.x, dy = c.y - q.y, dz = c.z - q.z;
                best3_insert(dx * dx + dy * dy + dz * dz, b0, b1, b2);
            }
        }
    }
    out[__float_as_uint(q.w)] = (b0 + b1 + b2) / 3.f;
}

struct KnnWs { size_t bbox, keys0, keys1, vals0, vals1, sorted, boxes, supers, sort_ws, total; };
static KnnWs knn_layout(int P) {
    KnnWs w; size_t o = 0;
    const size_t n = (size_t)(P > 0 ? P : 1), nb = (n + KNN_BOX - 1) / KNN_BOX;
    w.bbox = o; o += 256;
    w.keys0 = o; o += align_up(n * 8); w.keys1 = o; o += align_up(n * 8);
    w.vals0 = o; o += align_up(n * 4); w.vals1 = o; o += align_up(n * 4);
    w.sorted = o; o += align_up(n * 16);
    w.boxes = o; o += align_up(nb * 24);
    w.supers = o; o += align_up(((nb + KNN_SUPER - 1) / KNN_SUPER) * 24);
    w.sort_ws = o; o += align_up(sort_workspace_bytes((int64_t)n));
    w.total = o;
    return w;
}

size_t dist2_workspace_bytes(int P) { return P <= 2048 ? 256 : knn_layout(P).total; }

int launch_dist2(int P, const float *points, float *mean_dists, void *ws, size_t ws_bytes, cudaStream_t s) {
    if (P <= 0) return 0;
    if (P <= 2048) {
        LVDGS_PRE(s);
        dist2_bruteforce_kernel<<<ceil_div(P, KNN_THREADS), KNN_THREADS, 0, s>>>(P, points, mean_dists);
        LVDGS_LAUNCHED(s, "dist2_allpairs");
        return 0;
    }
    const KnnWs w = knn_layout(P);
    if (!ws || ws_bytes < w.total) { set_error("dist2: workspace too small (%zu < %zu)", ws_bytes, w.total); return 1; }
    char *b = (char *)ws;
    uint32_t *bbox = (uint32_t *)(b + w.bbox);
    uint64_t *k0 = (uint64_t *)(b + w.keys0), *k1 = (uint64_t *)(b + w.keys1);
    uint32_t *v0 = (uint32_t *)(b + w.vals0), *v1 = (uint32_t *)(b + w.vals1);
    float4 *sorted = (float4 *)(b + w.sorted);
    float *boxes = (float *)(b + w.boxes), *supers = (float *)(b + w.supers);
    const uint32_t init[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
    LVDGS_CHECK(cudaMemcpyAsync(bbox, init, sizeof init, cudaMemcpyHostToDevice, s));
    LVDGS_PRE(s);
    knn_bbox_kernel<<<min(148 * 4, ceil_div(P, KNN_THREADS)), KNN_THREADS, 0, s>>>(P, points, bbox);
    LVDGS_LAUNCHED(s, "knn_bbox");
    LVDGS_PRE(s);
    knn_morton_kernel<<<ceil_div(P, KNN_THREADS), KNN_THREADS, 0, s>>>(P, points, bbox, k0, v0);
    LVDGS_LAUNCHED(s, "knn_morton");
    int sel = 0;
    if (launch_sort_pairs(P, nullptr, k0, k1, v0, v1, 30, b + w.sort_ws, sort_workspace_bytes(P), nullptr, &sel, s)) return 1;
    const uint32_t *order = sel ? v1 : v0;
    const int nboxes = ceil_div(P, KNN_BOX);
    LVDGS_PRE(s);
    knn_boxes_kernel<<<nboxes, KNN_BOX, 0, s>>>(P, points, order, sorted, boxes);
    LVDGS_LAUNCHED(s, "knn_boxes");
    LVDGS_PRE(s);
    knn_super_kernel<<<ceil_div(ceil_div(nboxes, KNN_SUPER), KNN_THREADS), KNN_THREADS, 0, s>>>(nboxes, boxes, supers);
    LVDGS_LAUNCHED(s, "knn_super");
    LVDGS_PRE(s);
    knn_search_kernel<<<ceil_div(P, KNN_THREADS), KNN_THREADS, 0, s>>>(P, nboxes, sorted, boxes, supers, mean_dists);
    LVDGS_LAUNCHED(s, "knn_search");
    return 0;
}

}  // namespace lvdgs
