// radix_sort.cu -- K4: stable LSD radix sort of (u64 key, u32 value) pairs over key bits [0, end_bit),
// hand-written single-pass-per-digit "onesweep" (chained-scan with decoupled look-back), 8-bit digits.
//
// Replaces the cub::DeviceRadixSort::SortPairs call of the reference's rasterizer:
// keys are (tile_id << 32 | float_bits(depth)), values are Gaussian indices; the result must be the stable
// order, because equal (tile, depth) keys keep emission order.
//
// Per digit pass every block owns one tile of 4096 consecutive pairs: it ranks its keys (warp match-any
// ranking, stable), publishes its per-digit counts, resolves its global offsets by looking back over the
// predecessors' published counts, and scatters through shared memory so that global stores are runs of
// consecutive addresses per digit.  HBM traffic per pass: 12 B read + 12 B written per pair; one
// up-front histogram kernel reads the keys once for all passes.
#include "common.cuh"
#include <atomic>

namespace lvdgs {

#ifndef LVDGS_RS_THREADS
#define LVDGS_RS_THREADS 256
#endif
constexpr int RS_THREADS = LVDGS_RS_THREADS;
constexpr int RS_WARPS = RS_THREADS / 32;
#ifndef LVDGS_RS_ITEMS
#define LVDGS_RS_ITEMS 16
#endif
constexpr int RS_ITEMS = LVDGS_RS_ITEMS;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;   // 4096 pairs per block
constexpr int RS_BINS = 256;
constexpr int RS_MAX_PASSES = 8;
constexpr uint32_t LB_LOCAL = 1u << 30, LB_GLOBAL = 1u << 31, LB_MASK = (1u << 30) - 1;

struct SortWs {            // workspace header (zeroed before every sort)
    uint32_t hist[RS_MAX_PASSES][RS_BINS];
    uint32_t ticket[RS_MAX_PASSES];
    uint32_t pad[56];
};

size_t sort_workspace_bytes(int64_t n) {
    const size_t ntiles = (size_t)((n + RS_TILE - 1) / RS_TILE);
    return align_up(sizeof(SortWs)) + align_up(ntiles * RS_BINS * sizeof(uint32_t)) * RS_MAX_PASSES;
}

__device__ __forceinline__ uint32_t digit_of(uint64_t key, int shift, uint32_t mask) {
    return (uint32_t)(key >> shift) & mask;
}

// ---- histogram of every digit position in one read of the keys ----
// `n_dev` (optional): device-side element count, clamped to the capacity `n` the launch was sized for -- lets the
// forward launch the sort before the host has read the instance count back.
__device__ __forceinline__ int64_t rs_count(int64_t n_cap, const uint32_t *n_dev) {
    return n_dev ? min((int64_t)__ldg(n_dev), n_cap) : n_cap;
}

__global__ void __launch_bounds__(RS_THREADS) rs_histogram_kernel(const uint64_t *__restrict__ keys, int64_t n_cap,
                                                                  const uint32_t *__restrict__ n_dev, int passes,
                                                                  int end_bit, SortWs *ws) {
    const int64_t n = rs_count(n_cap, n_dev);
    __shared__ uint32_t h[RS_MAX_PASSES * RS_BINS];
    for (int k = threadIdx.x; k < passes * RS_BINS; k += RS_THREADS) h[k] = 0;
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * RS_THREADS;
    const int64_t n_round = (n + 31) / 32 * 32;    // keep warps converged for match_any
    for (int64_t i = (int64_t)blockIdx.x * RS_THREADS + threadIdx.x; i < n_round; i += stride) {
        const bool valid = i < n;
        const uint64_t key = valid ? __ldg(keys + i) : 0;
        for (int p = 0; p < passes; ++p) {
            const int shift = p * 8;
            const int nb = min(8, end_bit - shift);
            const uint32_t d = valid ? digit_of(key, shift, (1u << nb) - 1) : 0xffffffffu;
            const uint32_t peers = __match_any_sync(0xffffffffu, d);
            if (valid && (threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&h[p * RS_BINS + d], __popc(peers));
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < passes * RS_BINS; k += RS_THREADS) {
        const uint32_t c = h[k];
        if (c) atomicAdd(&ws->hist[k / RS_BINS][k % RS_BINS], c);
    }
}

__device__ __forceinline__ uint32_t rs_warp_incl_scan(uint32_t v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += n;
    }
    return v;
}

// exclusive scan over 256 threads (one value each)
__device__ __forceinline__ uint32_t rs_block_excl_scan(uint32_t v, uint32_t *warp_sums) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t incl = rs_warp_incl_scan(v);
    if (lane == 31) warp_sums[w] = incl;
    __syncthreads();
    if (w == 0) {
        uint32_t s = lane < RS_WARPS ? warp_sums[lane] : 0;
        s = rs_warp_incl_scan(s);
        if (lane < RS_WARPS) warp_sums[lane] = s;
    }
    __syncthreads();
    const uint32_t base = w ? warp_sums[w - 1] : 0;
    __syncthreads();
    return base + incl - v;
}

__global__ void __launch_bounds__(RS_THREADS) rs_scan_hist_kernel(SortWs *ws) {
    __shared__ uint32_t warp_sums[RS_WARPS];
    const uint32_t v = threadIdx.x < RS_BINS ? ws->hist[blockIdx.x][threadIdx.x] : 0u;
    const uint32_t e = rs_block_excl_scan(v, warp_sums);
    if (threadIdx.x < RS_BINS) ws->hist[blockIdx.x][threadIdx.x] = e;
}

#ifdef LVDGS_RS_TIMING     // experiment: per-phase clock() stamps of thread 0 of every block, [block][8]
__device__ long long g_rs_timing[4096 * 8];
#define RS_MARK(k) do { if (threadIdx.x == 0 && blockIdx.x < 4096) g_rs_timing[blockIdx.x * 8 + (k)] = clock64(); } while (0)
#else
#define RS_MARK(k) do { } while (0)
#endif

struct __align__(16) RsSmem {
    uint64_t keys[RS_TILE];                 // reused as uint32 vals[RS_TILE] in the second phase
    uint32_t cnt[RS_WARPS][RS_BINS];        // per-warp digit counters -> per-warp exclusive offsets
    uint32_t bin_start[RS_BINS];            // position of each digit's run inside the tile
    uint32_t goff[RS_BINS];                 // global index = tile position + goff[digit]
    uint32_t warp_sums[RS_WARPS];
    uint32_t tile;
    uint8_t dig[RS_TILE];
};

__global__ void __launch_bounds__(RS_THREADS) rs_onesweep_kernel(const uint64_t *__restrict__ keys_in, uint64_t *__restrict__ keys_out,
                                                                 const uint32_t *__restrict__ vals_in, uint32_t *__restrict__ vals_out,
                                                                 int64_t n_cap, const uint32_t *__restrict__ n_dev, int shift,
                                                                 int nbits, int pass, SortWs *ws, const uint32_t *__restrict__ hist_scanned,
                                                                 uint32_t *lookback) {
    // pre-computed histograms count ALL instances: if the launch capacity is too small (speculative forward whose hint
    // was exceeded) the scatter offsets would run past the buffers; that launch's output is discarded anyway -> retire
    if (n_dev && (int64_t)__ldg(n_dev) > n_cap) return;
    const int64_t n = rs_count(n_cap, n_dev);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RsSmem &sm = *reinterpret_cast<RsSmem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const uint32_t mask = (1u << nbits) - 1;

    if (tid == 0) sm.tile = atomicAdd(&ws->ticket[pass], 1u);
    for (int k = tid; k < RS_WARPS * RS_BINS; k += RS_THREADS) (&sm.cnt[0][0])[k] = 0;
    __syncthreads();
    RS_MARK(0);
    const uint32_t tile = sm.tile;
    const int64_t tile_base = (int64_t)tile * RS_TILE;
    if (tile_base >= n) return;          // launch was sized for the capacity; tiles past the real count retire at once
    const int n_valid = (int)min((int64_t)RS_TILE, n - tile_base);

    // ---- load (warp-striped: item k of lane l in warp w is element w*512 + k*32 + l of the tile) ----
    uint64_t key[RS_ITEMS];
    uint32_t val[RS_ITEMS];
#pragma unroll
    for (int k = 0; k < RS_ITEMS; ++k) {
        const int e = w * (32 * RS_ITEMS) + k * 32 + lane;
        key[k] = e < n_valid ? __ldg(keys_in + tile_base + e) : ~0ull;
        val[k] = e < n_valid ? __ldg(vals_in + tile_base + e) : 0u;
    }

    RS_MARK(1);
    // ---- stable rank inside the warp: items in order k = 0.., lanes ascending ----
    uint32_t rank[RS_ITEMS];
    const uint32_t lt_mask = (1u << lane) - 1;
#pragma unroll
    for (int k = 0; k < RS_ITEMS; ++k) {
        const uint32_t d = digit_of(key[k], shift, mask);
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (lane == leader) {
            old = sm.cnt[w][d];
            sm.cnt[w][d] = old + __popc(peers);
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[k] = old + __popc(peers & lt_mask);
        __syncwarp();
    }
    __syncthreads();
    RS_MARK(2);

    // ---- thread d (< 256) owns digit d: warp-exclusive offsets, tile total, look-back ----
    uint32_t total = 0;
    if (tid < RS_BINS) {
#pragma unroll
        for (int ww = 0; ww < RS_WARPS; ++ww) {
            const uint32_t c = sm.cnt[ww][tid];
            sm.cnt[ww][tid] = total;
            total += c;
        }
    }
    const uint32_t bstart = rs_block_excl_scan(total, sm.warp_sums);
    if (tid < RS_BINS) {
        sm.bin_start[tid] = bstart;
        volatile uint32_t *lb = lookback;
        uint32_t excl = 0;
#ifdef LVDGS_RS_NOLOOKBACK      // timing experiment only: results are wrong
        if (true) {
#else
        if (tile == 0) {
#endif
            lb[(size_t)tile * RS_BINS + tid] = total | LB_GLOBAL;
        } else {
            lb[(size_t)tile * RS_BINS + tid] = total | LB_LOCAL;
            int64_t t = (int64_t)tile - 1;
            while (true) {
                const uint32_t v = lb[(size_t)t * RS_BINS + tid];
                if (v & (LB_LOCAL | LB_GLOBAL)) {
                    excl += v & LB_MASK;
                    if (v & LB_GLOBAL) break;
                    --t;
                }
            }
            lb[(size_t)tile * RS_BINS + tid] = (excl + total) | LB_GLOBAL;
        }
        sm.goff[tid] = hist_scanned[pass * RS_BINS + tid] + excl - bstart;
    }
    __syncthreads();
    RS_MARK(3);

    // ---- scatter keys through shared memory, then coalesced runs to global ----
    uint32_t pos[RS_ITEMS];
#pragma unroll
    for (int k = 0; k < RS_ITEMS; ++k) {
        const uint32_t d = digit_of(key[k], shift, mask);
        pos[k] = sm.bin_start[d] + sm.cnt[w][d] + rank[k];
        sm.keys[pos[k]] = key[k];
        sm.dig[pos[k]] = (uint8_t)d;
    }
    __syncthreads();
    RS_MARK(4);
#pragma unroll
    for (int k = 0; k < RS_ITEMS; ++k) {
        const int s = k * RS_THREADS + tid;
        if (s < n_valid) keys_out[(uint32_t)(s + sm.goff[sm.dig[s]])] = sm.keys[s];
    }
    __syncthreads();
    RS_MARK(5);
    uint32_t *svals = reinterpret_cast<uint32_t *>(sm.keys);
#pragma unroll
    for (int k = 0; k < RS_ITEMS; ++k) svals[pos[k]] = val[k];
    __syncthreads();
    RS_MARK(6);
#pragma unroll
    for (int k = 0; k < RS_ITEMS; ++k) {
        const int s = k * RS_THREADS + tid;
        if (s < n_valid) vals_out[(uint32_t)(s + sm.goff[sm.dig[s]])] = svals[s];
    }
    RS_MARK(7);
}

int launch_sort_pairs(int64_t n, const uint32_t *n_dev, uint64_t *keys0, uint64_t *keys1, uint32_t *vals0,
                      uint32_t *vals1, int end_bit, void *ws_raw, size_t ws_bytes, const uint32_t *pre_hist, int *selector,
                      cudaStream_t s) {
    if (selector) *selector = 0;
    if (n <= 0) return 0;
    if (end_bit < 1 || end_bit > 64) { set_error("sort: end_bit %d out of range", end_bit); return 1; }
    if (n >= (1ll << 30)) { set_error("sort: n=%lld exceeds 2^30", (long long)n); return 1; }
    if (ws_bytes < sort_workspace_bytes(n)) { set_error("sort: workspace too small"); return 1; }
    const int passes = (end_bit + 7) / 8;
    const int ntiles = ceil_div(n, RS_TILE);
    SortWs *ws = reinterpret_cast<SortWs *>(ws_raw);
    unsigned char *lb_base = reinterpret_cast<unsigned char *>(ws_raw) + align_up(sizeof(SortWs));
    const size_t lb_stride = align_up((size_t)ntiles * RS_BINS * sizeof(uint32_t));
    LVDGS_CHECK(cudaMemsetAsync(ws_raw, 0, align_up(sizeof(SortWs)) + lb_stride * passes, s));
    static std::atomic<bool> attr_set[MAX_DEVICES];          // per device (see common.cuh)
    const int dev = current_device();
    if (!attr_set[dev].load(std::memory_order_acquire)) {
        LVDGS_CHECK(cudaFuncSetAttribute(rs_onesweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RsSmem)));
        attr_set[dev].store(true, std::memory_order_release);
    }
    const uint32_t *hist_scanned = pre_hist;
    if (!pre_hist) {
        const int hist_blocks = (int)min((int64_t)148 * 8, (n + RS_THREADS * 8 - 1) / (RS_THREADS * 8));
        LVDGS_PRE(s);
        rs_histogram_kernel<<<hist_blocks, RS_THREADS, 0, s>>>(keys0, n, n_dev, passes, end_bit, ws);
        LVDGS_LAUNCHED(s, "sort_histogram");
        LVDGS_PRE(s);
        rs_scan_hist_kernel<<<passes, RS_THREADS, 0, s>>>(ws);
        LVDGS_LAUNCHED(s, "sort_scan_hist");
        hist_scanned = &ws->hist[0][0];
    }
    uint64_t *kin = keys0, *kout = keys1;
    uint32_t *vin = vals0, *vout = vals1;
    for (int p = 0; p < passes; ++p) {
        const int shift = p * 8, nb = min(8, end_bit - shift);
        LVDGS_PRE(s);
        rs_onesweep_kernel<<<ntiles, RS_THREADS, sizeof(RsSmem), s>>>(kin, kout, vin, vout, n, n_dev, shift, nb, p, ws, hist_scanned,
                                                                      reinterpret_cast<uint32_t *>(lb_base + lb_stride * p));
        LVDGS_LAUNCHED(s, "sort_onesweep");
        uint64_t *tk = kin; kin = kout; kout = tk;
        uint32_t *tv = vin; vin = vout; vout = tv;
    }
    if (selector) *selector = passes & 1;
    return 0;
}

#ifdef LVDGS_RS_TIMING
extern "C" int lvdgs_debug_sort_timing(long long *host, int n) {
    return (int)cudaMemcpyFromSymbol(host, g_rs_timing, sizeof(long long) * n);
}
#endif

}  // namespace lvdgs
