// tile_sort.cu -- K4 as a segmented sort: one CTA sorts one tile's instance list in shared memory.
//
// The reference sorts all R (tile | depth) keys with one global stable radix sort (cub::DeviceRadixSort, SURVEY.md K4:
// 6 passes x 24 B per instance through HBM, and at R ~ 1-2 M every pass is one latency-bound wave).  Here the tile
// ranges are known BEFORE any key exists (binning_prep integrates the per-tile counts), so the key emission drops every
// instance straight into its tile's segment (slot = atomicAdd on the tile's cursor) and what is left is ~2000
// independent sorts of ~1000 elements -- shared-memory work with no HBM round trips.
//
// Bit-exactness: the stable global sort orders a tile's instances by depth bits, ties by emission order = ascending
// Gaussian index (a Gaussian appears at most once per tile).  That is the total order of the 64-bit word
// (depth bits << 32 | Gaussian index), whose sorted sequence is unique -- so ANY sorting network yields the
// reference's list, regardless of the (non-deterministic) slot order the atomics produced.
//
// Network: bitonic sort in its all-ascending form (first step of every merge level compares i with its mirror image
// inside the 2^k block, the remaining steps compare i with i + j).  Because every comparator moves the smaller word to
// the lower index, elements beyond n behave like +inf without being stored: comparators whose upper index is >= n are
// skipped, and with the index maps below the live comparators of a stage are a PREFIX [0, cnt) -- the work is
// proportional to n, not to the next power of two.
// Three size classes (tile_order lists the tiles by decreasing length, so each class is a contiguous run of it):
//   n < 2048        one 256-thread CTA per tile; sorts 32-bit stand-ins (quantised depth | slot), fixes up ties;
//   2048 .. 8191    the same with 13 slot bits, persistent 512-thread CTAs;
//   >= 8192         64-bit network, persistent 1024-thread CTAs, 16384 words of shared memory; longer lists run the wide
//                   stages in place in global memory (L2) and the narrow ones chunk by chunk in shared memory.
// The two upper classes are launched only when the previous frame's longest list calls for them (api.cu).
#include "common.cuh"
#include <atomic>

#ifndef LVDGS_TS_Q32
#define LVDGS_TS_Q32 1        // short lists: 32-bit stand-ins + fix-up (0: the 64-bit network directly)
#endif

namespace lvdgs {

__device__ __forceinline__ uint64_t lds_u64(uint32_t a) {
    uint64_t v;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u64(uint32_t a, uint64_t v) {
    asm volatile("st.shared.u64 [%0], %1;" :: "r"(a), "l"(v) : "memory");
}

// ---- register-blocked network --------------------------------------------------------------------------------------
// A plain shared-memory bitonic sort is bound by the LSU pipe (4 accesses per comparator, 66 stages at n = 2048, 2-4-way
// bank conflicts at small strides).  Here a thread takes EIGHT words into registers and runs up to three consecutive
// stages on them before writing back, so a round trip through shared memory serves 12 comparators instead of 4:
//   * levels 2, 4, 8            : 8 consecutive words, all six stages in one round trip;
//   * head of level K >= 16     : 4 words r + S m of the lower half of a K-block and their 4 mirror images (S = K/8):
//                                 the flip stage and the stages j = K/4, K/8;
//   * rest of the level         : words i0 + S m (m = 0..7): the stages j = 4S, 2S, S, three at a time down to j = 1.
// Word i lives at i + (i >> 4) (one pad word per 16): every access pattern above is conflict-free (two wavefronts per
// 64-bit warp access) AND the eight addresses of a thread are its first address plus compile-time constants
// (d + (d >> 4) for d = S m: the low four bits never carry, see the index maps), so a round trip is one address
// computation + 8 LDS + 8 STS with immediate offsets.  Words [n, next power of two) are physically filled with TS_PAD,
// which removes every bounds test from the loads and stores (a comparator never moves a pad word below n).
// 24 round trips for n = 2048 instead of 66 conflicted stages; the kernel is bound by the ALU pipe (6 ALU instructions
// per 64-bit comparator: sm_100a has no 64-bit min/max, and DMNMX on the words read as doubles does not exist either).
constexpr uint64_t TS_PAD = ~0ull;
constexpr int ts_phys(int i) { return i + (i >> 4); }
constexpr size_t ts_smem_bytes(int cap) { return (size_t)ts_phys(cap) * sizeof(uint64_t); }
__device__ __forceinline__ uint32_t ts_addr(uint32_t a_s, int i) { return a_s + (uint32_t)((i + (i >> 4)) << 3); }
__device__ __forceinline__ void ce(uint64_t &a, uint64_t &b) {        // a sits at the lower index
    const bool sw = a > b;
    const uint64_t lo = sw ? b : a, hi = sw ? a : b;
    a = lo; b = hi;
}

template <int THREADS>
__device__ __forceinline__ void ts_levels_2_4_8(uint32_t a_s, int n, int tid) {
    __syncthreads();
    for (int g = tid; g * 8 < n; g += THREADS) {
        const uint32_t a0 = ts_addr(a_s, 8 * g);
        uint64_t v[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) v[m] = lds_u64(a0 + 8 * m);
        ce(v[0], v[1]); ce(v[2], v[3]); ce(v[4], v[5]); ce(v[6], v[7]);
        ce(v[0], v[3]); ce(v[1], v[2]); ce(v[4], v[7]); ce(v[5], v[6]);
        ce(v[0], v[1]); ce(v[2], v[3]); ce(v[4], v[5]); ce(v[6], v[7]);
        ce(v[0], v[7]); ce(v[1], v[6]); ce(v[2], v[5]); ce(v[3], v[4]);
        ce(v[0], v[2]); ce(v[1], v[3]); ce(v[4], v[6]); ce(v[5], v[7]);
        ce(v[0], v[1]); ce(v[2], v[3]); ce(v[4], v[5]); ce(v[6], v[7]);
#pragma unroll
        for (int m = 0; m < 8; ++m) sts_u64(a0 + 8 * m, v[m]);
    }
}

// flip stage of level K and the stages j = K/4, K/8
template <int THREADS, int K>
__device__ __forceinline__ void ts_flip_group(uint32_t a_s, int n, int tid) {
    constexpr int S = K / 8;
    const int G = ((n + K - 1) / K) * S;
    __syncthreads();
    for (int g = tid; g < G; g += THREADS) {
        const int r = g & (S - 1), base = (g / S) * K;
        const int il = base + r, iu = base + K - 1 - r;
        if (il >= n) continue;
        const uint32_t al = ts_addr(a_s, il), au = ts_addr(a_s, iu);
        uint64_t lo[4], up[4];
#pragma unroll
        for (int m = 0; m < 4; ++m) { lo[m] = lds_u64(al + 8 * ts_phys(S * m)); up[m] = lds_u64(au - 8 * ts_phys(S * m)); }
#pragma unroll
        for (int m = 0; m < 4; ++m) ce(lo[m], up[m]);
        ce(lo[0], lo[2]); ce(lo[1], lo[3]); ce(up[2], up[0]); ce(up[3], up[1]);
        ce(lo[0], lo[1]); ce(lo[2], lo[3]); ce(up[1], up[0]); ce(up[3], up[2]);
#pragma unroll
        for (int m = 0; m < 4; ++m) { sts_u64(al + 8 * ts_phys(S * m), lo[m]); sts_u64(au - 8 * ts_phys(S * m), up[m]); }
    }
}

// the last CNT of the stages j = 4S, 2S, S
template <int THREADS, int S, int CNT>
__device__ __forceinline__ void ts_j_group(uint32_t a_s, int n, int tid) {
    const int G = ((n + 8 * S - 1) / (8 * S)) * S;
    __syncthreads();
    for (int g = tid; g < G; g += THREADS) {
        const int i0 = (g / S) * (8 * S) + (g & (S - 1));
        if (i0 + S >= n) continue;             // fewer than two live words: nothing to compare
        const uint32_t a0 = ts_addr(a_s, i0);
        uint64_t v[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) v[m] = lds_u64(a0 + 8 * ts_phys(S * m));
        if (CNT >= 3) { ce(v[0], v[4]); ce(v[1], v[5]); ce(v[2], v[6]); ce(v[3], v[7]); }
        if (CNT >= 2) { ce(v[0], v[2]); ce(v[1], v[3]); ce(v[4], v[6]); ce(v[5], v[7]); }
        ce(v[0], v[1]); ce(v[2], v[3]); ce(v[4], v[5]); ce(v[6], v[7]);
#pragma unroll
        for (int m = 0; m < 8; ++m) sts_u64(a0 + 8 * ts_phys(S * m), v[m]);
    }
}

constexpr int ts_ilog2(int x) { return x <= 1 ? 0 : 1 + ts_ilog2(x / 2); }

// stages j = J, J/2, ... 1
template <int THREADS, int J>
__device__ __forceinline__ void ts_j_stages_from(uint32_t a_s, int n, int tid) {
    if constexpr (J >= 1) {
        constexpr int CNT = ts_ilog2(J) + 1 >= 3 ? 3 : ts_ilog2(J) + 1;
        constexpr int S = J >> (CNT - 1);
        if (S < n) ts_j_group<THREADS, S, CNT>(a_s, n, tid);
        ts_j_stages_from<THREADS, S / 2>(a_s, n, tid);
    }
}

// merge levels K, 2K, ... CAP (K >= 16) on the n (<= CAP) words at shared address a_s
template <int THREADS, int K, int CAP>
__device__ __forceinline__ void ts_levels_from(uint32_t a_s, int n, int tid) {
    if constexpr (K <= CAP) {
        if (K / 2 < n) {              // otherwise the data is already one sorted run
            ts_flip_group<THREADS, K>(a_s, n, tid);
            ts_j_stages_from<THREADS, K / 16>(a_s, n, tid);
            ts_levels_from<THREADS, K * 2, CAP>(a_s, n, tid);
        }
    }
}

// words [n, upto) := TS_PAD; upto = 0: the next power of two >= max(n, 8), which bounds every index a full sort touches
template <int THREADS>
__device__ __forceinline__ void ts_fill_pad(uint32_t a_s, int n, int tid, int upto = 0) {
    const int npad = upto ? upto : (n <= 8 ? 8 : 1 << (32 - __clz(n - 1)));
    for (int i = n + tid; i < npad; i += THREADS) sts_u64(ts_addr(a_s, i), TS_PAD);
}

// full sort of n <= CAP words; the caller has written them (through ts_addr) and must barrier before reading them back
template <int THREADS, int CAP>
__device__ __forceinline__ void ts_sort_smem(uint32_t a_s, int n, int tid) {
    ts_fill_pad<THREADS>(a_s, n, tid);
    ts_levels_2_4_8<THREADS>(a_s, n, tid);
    ts_levels_from<THREADS, 16, CAP>(a_s, n, tid);
}

// ---- 32-bit network for the short lists -------------------------------------------------------------------------
// A 64-bit comparator is 6 ALU instructions; a 32-bit one is two (VIMNMX min / max).  For a list of n < 2048 instances
// the sort therefore runs on 32-bit stand-ins  q << 11 | slot,  q = (depth bits - list minimum) >> shift  quantised to 21
// bits (monotone, so the order by (q, slot) is the true order except INSIDE runs of equal q) and slot = the instance's
// position in the unsorted segment (unique).  A fix-up pass then orders every run of equal q by the full 64-bit words
// (runs are a handful of instances unless many depths coincide to within 2^-21 of the list's depth range; runs beyond
// TS_MAXRUN send the whole list through the 64-bit network instead -- still exact, just slower).
// Two stand-ins share one 64-bit shared-memory word (even element = low half), which keeps the word-level index maps,
// the conflict-free layout and the immediate-offset addressing of the 64-bit network: a stage with stride j >= 2 is the
// word-level stage j/2 applied to both halves, the stride-1 stage compares the halves of a word, and a flip compares a
// word's halves with the opposite halves of its mirror word.  A thread holds 8 words = 16 elements.
__device__ __forceinline__ void ce32(uint32_t &a, uint32_t &b) {
    const uint32_t lo = min(a, b), hi = max(a, b);
    a = lo; b = hi;
}
struct PairWords {
    uint32_t e[8], o[8];      // even / odd element of 8 words
    __device__ __forceinline__ void load(int m, uint32_t addr) { const uint64_t w = lds_u64(addr); e[m] = (uint32_t)w; o[m] = (uint32_t)(w >> 32); }
    __device__ __forceinline__ void store(int m, uint32_t addr) const { sts_u64(addr, (uint64_t)o[m] << 32 | e[m]); }
    __device__ __forceinline__ void cew(int i, int j) { ce32(e[i], e[j]); ce32(o[i], o[j]); }         // word i below word j
    __device__ __forceinline__ void inword(int i) { ce32(e[i], o[i]); }
    __device__ __forceinline__ void flip(int a, int b) { ce32(e[a], o[b]); ce32(o[a], e[b]); }         // word a below its mirror b
    __device__ __forceinline__ void inword_all() {
#pragma unroll
        for (int i = 0; i < 8; ++i) inword(i);
    }
};

// element levels 2, 4, 8, 16 on 8 consecutive words
template <int THREADS>
__device__ __forceinline__ void tp_levels_2_to_16(uint32_t a_s, int nw, int tid) {
    __syncthreads();
    for (int g = tid; g * 8 < nw; g += THREADS) {
        const uint32_t a0 = ts_addr(a_s, 8 * g);
        PairWords w;
#pragma unroll
        for (int m = 0; m < 8; ++m) w.load(m, a0 + 8 * m);
        w.inword_all();
        w.flip(0, 1); w.flip(2, 3); w.flip(4, 5); w.flip(6, 7);
        w.inword_all();
        w.flip(0, 3); w.flip(1, 2); w.flip(4, 7); w.flip(5, 6);
        w.cew(0, 1); w.cew(2, 3); w.cew(4, 5); w.cew(6, 7);
        w.inword_all();
        w.flip(0, 7); w.flip(1, 6); w.flip(2, 5); w.flip(3, 4);
        w.cew(0, 2); w.cew(1, 3); w.cew(4, 6); w.cew(5, 7);
        w.cew(0, 1); w.cew(2, 3); w.cew(4, 5); w.cew(6, 7);
        w.inword_all();
#pragma unroll
        for (int m = 0; m < 8; ++m) w.store(m, a0 + 8 * m);
    }
}

// word-level flip of level KW (element level 2 KW) and the word stages KW/4, KW/8
template <int THREADS, int KW>
__device__ __forceinline__ void tp_flip_group(uint32_t a_s, int nw, int tid) {
    constexpr int S = KW / 8;
    const int G = ((nw + KW - 1) / KW) * S;
    __syncthreads();
    for (int g = tid; g < G; g += THREADS) {
        const int r = g & (S - 1), base = (g / S) * KW;
        const int il = base + r, iu = base + KW - 1 - r;
        if (il >= nw) continue;
        const uint32_t al = ts_addr(a_s, il), au = ts_addr(a_s, iu);
        PairWords w;          // 0..3: lower words, ascending; 4..7: mirror words, DESCENDING
#pragma unroll
        for (int m = 0; m < 4; ++m) { w.load(m, al + 8 * ts_phys(S * m)); w.load(4 + m, au - 8 * ts_phys(S * m)); }
#pragma unroll
        for (int m = 0; m < 4; ++m) w.flip(m, 4 + m);
        w.cew(0, 2); w.cew(1, 3); w.cew(6, 4); w.cew(7, 5);
        w.cew(0, 1); w.cew(2, 3); w.cew(5, 4); w.cew(7, 6);
#pragma unroll
        for (int m = 0; m < 4; ++m) { w.store(m, al + 8 * ts_phys(S * m)); w.store(4 + m, au - 8 * ts_phys(S * m)); }
    }
}

// the last CNT of the word stages 4S, 2S, S; S == 1 also ends the level with the in-word (element stride 1) stage
template <int THREADS, int S, int CNT>
__device__ __forceinline__ void tp_j_group(uint32_t a_s, int nw, int tid) {
    const int G = ((nw + 8 * S - 1) / (8 * S)) * S;
    __syncthreads();
    for (int g = tid; g < G; g += THREADS) {
        const int i0 = (g / S) * (8 * S) + (g & (S - 1));
        if (i0 >= nw) continue;
        const uint32_t a0 = ts_addr(a_s, i0);
        PairWords w;
#pragma unroll
        for (int m = 0; m < 8; ++m) w.load(m, a0 + 8 * ts_phys(S * m));
        if (CNT >= 3) { w.cew(0, 4); w.cew(1, 5); w.cew(2, 6); w.cew(3, 7); }
        if (CNT >= 2) { w.cew(0, 2); w.cew(1, 3); w.cew(4, 6); w.cew(5, 7); }
        w.cew(0, 1); w.cew(2, 3); w.cew(4, 5); w.cew(6, 7);
        if (S == 1) w.inword_all();
#pragma unroll
        for (int m = 0; m < 8; ++m) w.store(m, a0 + 8 * ts_phys(S * m));
    }
}

template <int THREADS, int J>
__device__ __forceinline__ void tp_j_stages_from(uint32_t a_s, int nw, int tid) {
    if constexpr (J >= 1) {
        constexpr int CNT = ts_ilog2(J) + 1 >= 3 ? 3 : ts_ilog2(J) + 1;
        constexpr int S = J >> (CNT - 1);
        if (S < nw || S == 1) tp_j_group<THREADS, S, CNT>(a_s, nw, tid);
        tp_j_stages_from<THREADS, S / 2>(a_s, nw, tid);
    }
}

template <int THREADS, int KW, int CAPW>
__device__ __forceinline__ void tp_levels_from(uint32_t a_s, int nw, int tid) {
    if constexpr (KW <= CAPW) {
        if (KW / 2 < nw) {
            tp_flip_group<THREADS, KW>(a_s, nw, tid);
            tp_j_stages_from<THREADS, KW / 16>(a_s, nw, tid);
            tp_levels_from<THREADS, KW * 2, CAPW>(a_s, nw, tid);
        }
    }
}

constexpr int TS_MAXRUN = 16;

// Sorts the n <= CAP = 2^TS_SLOT_BITS words of one tile.  s_words: ts_phys(CAP) words (the unsorted 64-bit words, by slot),
// s_pairs: ts_phys(CAP / 2) pair words, s_red: 2 * THREADS / 32 + 1 words of scratch.
template <int THREADS, int CAP>
__device__ __forceinline__ void tile_sort_one_q32(int tile, uint2 range, const uint64_t *seg, uint64_t *__restrict__ keys_out,
                                                  uint32_t *__restrict__ vals_out, uint64_t *s_words, uint64_t *s_pairs,
                                                  uint32_t *s_red) {
    constexpr int TS_SLOT_BITS = ts_ilog2(CAP), TS_Q_BITS = 32 - TS_SLOT_BITS;
    static_assert(CAP == (1 << TS_SLOT_BITS), "CAP must be a power of two");
    const int n = (int)(range.y - range.x), nw = (n + 1) >> 1, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t a_w = smem_u32(s_words), a_p = smem_u32(s_pairs);
    const uint64_t *g = seg + range.x;
    const uint64_t tile_hi = (uint64_t)(uint32_t)tile << 32;
    // (1) stage the words, find the range of their depth bits
    uint32_t dmin = 0xffffffffu, dmax = 0u;
    for (int i = tid; i < n; i += THREADS) {
        const uint64_t w = g[i];
        sts_u64(ts_addr(a_w, i), w);
        const uint32_t d = (uint32_t)(w >> 32);
        dmin = min(dmin, d); dmax = max(dmax, d);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        dmin = min(dmin, __shfl_xor_sync(0xffffffffu, dmin, d));
        dmax = max(dmax, __shfl_xor_sync(0xffffffffu, dmax, d));
    }
    if (lane == 0) { s_red[2 * warp] = dmin; s_red[2 * warp + 1] = dmax; }
    if (tid == 0) s_red[2 * (THREADS / 32)] = 0;          // "a run was too long" flag
    __syncthreads();
#pragma unroll
    for (int k = 0; k < THREADS / 32; ++k) { dmin = min(dmin, s_red[2 * k]); dmax = max(dmax, s_red[2 * k + 1]); }
    const int span_bits = 32 - __clz(dmax - dmin);         // 0 when all depths coincide
    const int shift = max(0, span_bits - TS_Q_BITS);
    // (2) the 32-bit stand-ins, two per pair word; elements beyond n and words beyond nw are all-ones
    {
        const int npw = nw <= 8 ? 8 : 1 << (32 - __clz(nw - 1));
        for (int w = tid; w < npw; w += THREADS) {
            uint32_t k0 = 0xffffffffu, k1 = 0xffffffffu;
            if (2 * w < n) k0 = (((uint32_t)(lds_u64(ts_addr(a_w, 2 * w)) >> 32) - dmin) >> shift) << TS_SLOT_BITS | (uint32_t)(2 * w);
            if (2 * w + 1 < n) k1 = (((uint32_t)(lds_u64(ts_addr(a_w, 2 * w + 1)) >> 32) - dmin) >> shift) << TS_SLOT_BITS | (uint32_t)(2 * w + 1);
            sts_u64(ts_addr(a_p, w), (uint64_t)k1 << 32 | k0);
        }
    }
    // (3) sort them
    tp_levels_2_to_16<THREADS>(a_p, nw, tid);
    tp_levels_from<THREADS, 16, CAP / 2>(a_p, nw, tid);
    __syncthreads();
    // (4) fix-up: a stand-in alone in its run is in its final place; inside a run the full words decide
    auto key_at = [&](int p) -> uint32_t {
        uint32_t v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(ts_addr(a_p, p >> 1) + 4u * (uint32_t)(p & 1)) : "memory");
        return v;
    };
    bool too_long = false;
    for (int p = tid; p < n; p += THREADS) {
        const uint32_t key = key_at(p), q = key >> TS_SLOT_BITS;
        const uint64_t my = lds_u64(ts_addr(a_w, (int)(key & ((1u << TS_SLOT_BITS) - 1u))));
        int pos = p;
        const bool pe = p > 0 && (key_at(p - 1) >> TS_SLOT_BITS) == q, ne = p + 1 < n && (key_at(p + 1) >> TS_SLOT_BITS) == q;
        if (pe || ne) {
            int s = p, e = p + 1;
            while (s > 0 && p - s < TS_MAXRUN && (key_at(s - 1) >> TS_SLOT_BITS) == q) --s;
            while (e < n && e - p < TS_MAXRUN && (key_at(e) >> TS_SLOT_BITS) == q) ++e;
            if (p - s >= TS_MAXRUN || e - p >= TS_MAXRUN) too_long = true;
            int rank = 0;
            for (int r = s; r < e; ++r)
                rank += lds_u64(ts_addr(a_w, (int)(key_at(r) & ((1u << TS_SLOT_BITS) - 1u)))) < my ? 1 : 0;
            pos = s + rank;
        }
        keys_out[range.x + pos] = tile_hi | (my >> 32);
        vals_out[range.x + pos] = (uint32_t)my;
    }
    if (too_long) s_red[2 * (THREADS / 32)] = 1;
    __syncthreads();
    if (s_red[2 * (THREADS / 32)]) {        // block-uniform: the 64-bit network on the staged words
        ts_sort_smem<THREADS, CAP>(a_w, n, tid);
        __syncthreads();
        for (int i = tid; i < n; i += THREADS) {
            const uint64_t w = lds_u64(ts_addr(a_w, i));
            keys_out[range.x + i] = tile_hi | (w >> 32);
            vals_out[range.x + i] = (uint32_t)w;
        }
    }
}

__device__ __forceinline__ void ce_global(uint64_t *g, int lo, int hi) {
    const uint64_t a = g[lo], b = g[hi];
    if (a > b) { g[lo] = b; g[hi] = a; }
}

// Sorts the segments of the tiles whose length n satisfies n_lo <= n < n_hi (one launch per size class).  tile_order
// lists the tiles by decreasing length class (quarter-octave buckets, so a power-of-two boundary never splits a bucket):
// the long-list class runs as a few persistent CTAs that walk the head of tile_order and stop at the first short list;
// the short-list class has one CTA per tile.  seg: (depth bits << 32 | Gaussian) in slot order (sorted in place when
// n > CAP).
template <int THREADS, int CAP, bool CHUNKED>
__device__ __forceinline__ void tile_sort_one(int tile, uint2 range, uint64_t *seg, uint64_t *__restrict__ keys_out,
                                              uint32_t *__restrict__ vals_out, uint64_t *s_words) {
    const uint32_t n32 = range.y - range.x;
    const int n = (int)n32, tid = threadIdx.x;
    const uint32_t a_s = smem_u32(s_words);
    uint64_t *g = seg + range.x;
    const uint64_t tile_hi = (uint64_t)(uint32_t)tile << 32;
    if (!CHUNKED || n <= CAP) {
        for (int i = tid; i < n; i += THREADS) sts_u64(ts_addr(a_s, i), g[i]);
        ts_sort_smem<THREADS, CAP>(a_s, n, tid);
        __syncthreads();
        for (int i = tid; i < n; i += THREADS) {
            const uint64_t w = lds_u64(ts_addr(a_s, i));
            keys_out[range.x + i] = tile_hi | (w >> 32);
            vals_out[range.x + i] = (uint32_t)w;
        }
        return;
    }
    if constexpr (CHUNKED) {
        // phase 1: every CAP-chunk becomes a sorted run
        for (int base = 0; base < n; base += CAP) {
            const int cn = min(CAP, n - base);
            __syncthreads();
            for (int i = tid; i < cn; i += THREADS) sts_u64(ts_addr(a_s, i), g[base + i]);
            ts_sort_smem<THREADS, CAP>(a_s, cn, tid);
            __syncthreads();
            for (int i = tid; i < cn; i += THREADS) g[base + i] = lds_u64(ts_addr(a_s, i));
        }
        // phase 2: merge levels above CAP; wide stages in global memory, stages below CAP per chunk in shared memory
        for (int64_t k = 2 * (int64_t)CAP; (k >> 1) < n; k <<= 1) {
            __syncthreads();
            {
                const int hk = (int)(k >> 1);
                const int cnt = (int)(n / k) * hk + max(0, (int)(n % k) - hk);
                for (int c = tid; c < cnt; c += THREADS) {
                    const int r = c & (hk - 1), mid = ((c - r) << 1) + hk;
                    ce_global(g, mid - 1 - r, mid + r);
                }
            }
            for (int j = (int)(k >> 2); j >= CAP; j >>= 1) {
                __syncthreads();
                const int cnt = (n / (2 * j)) * j + max(0, n % (2 * j) - j);
                for (int c = tid; c < cnt; c += THREADS) {
                    const int lo = c + (c & ~(j - 1));
                    ce_global(g, lo, lo + j);
                }
            }
            for (int base = 0; base < n; base += CAP) {
                const int cn = min(CAP, n - base);
                __syncthreads();
                for (int i = tid; i < cn; i += THREADS) sts_u64(ts_addr(a_s, i), g[base + i]);
                ts_fill_pad<THREADS>(a_s, cn, tid, CAP);     // the wide groups of a partial chunk reach up to CAP
                ts_j_stages_from<THREADS, CAP / 2>(a_s, cn, tid);
                __syncthreads();
                for (int i = tid; i < cn; i += THREADS) g[base + i] = lds_u64(ts_addr(a_s, i));
            }
        }
        __syncthreads();
        for (int i = tid; i < n; i += THREADS) {
            const uint64_t w = g[i];
            keys_out[range.x + i] = tile_hi | (w >> 32);
            vals_out[range.x + i] = (uint32_t)w;
        }
    }
}

// Lists of n_hi or more instances belong to the long-list launch.  passthrough: there is no such launch (the host
// expected no long list) -- the list is then handed on UNSORTED, so that the speculatively queued blend reads valid
// Gaussian indices; the host sees the longest list next to R and re-runs the tail with the long-list class.
template <int THREADS, int CAP>
__global__ void __launch_bounds__(THREADS) tile_sort_short_kernel(uint32_t n_hi, int passthrough, uint32_t capacity, const uint32_t *__restrict__ n_dev,
                                                                  const uint2 *__restrict__ ranges, const uint32_t *__restrict__ tile_order,
                                                                  uint64_t *seg, uint64_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out) {
    __shared__ uint64_t s_words[ts_phys(CAP)];
#if LVDGS_TS_Q32
    __shared__ uint64_t s_pairs[ts_phys(CAP / 2)];
    __shared__ uint32_t s_red[2 * (THREADS / 32) + 1];
#endif
    pdl_wait();                                  // launched behind the key emission (or the middle class): programmatic dependent launch
    const int tile = (int)__ldg(tile_order + blockIdx.x);
    const uint2 range = ranges[tile];
    const uint32_t n = range.y - range.x;
    if (n == 0) return;
    if (n_dev && __ldg(n_dev) > capacity) return;       // speculative launch with too small an arena: the host re-runs
    if (n >= n_hi) {
        if (passthrough)
            for (uint32_t i = threadIdx.x; i < n; i += THREADS) vals_out[range.x + i] = (uint32_t)seg[range.x + i];
        return;
    }
#if LVDGS_TS_Q32
    tile_sort_one_q32<THREADS, CAP>(tile, range, seg, keys_out, vals_out, s_words, s_pairs, s_red);
#else
    tile_sort_one<THREADS, CAP, false>(tile, range, seg, keys_out, vals_out, s_words);
#endif
}

// Middle class: lists of n_lo <= n < n_hi instances, 32-bit stand-ins like the short class but with CAP = 8192 (13 slot bits,
// 19 depth bits), persistent CTAs walking the head of tile_order (longer lists first, so they skip what the long class
// owns and stop at the first short list).
template <int THREADS, int CAP>
__global__ void __launch_bounds__(THREADS) tile_sort_mid_kernel(uint32_t n_lo, uint32_t n_hi, int tiles, uint32_t capacity, const uint32_t *__restrict__ n_dev,
                                                                const uint2 *__restrict__ ranges, const uint32_t *__restrict__ tile_order,
                                                                uint64_t *seg, uint64_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out) {
    extern __shared__ uint64_t s_dyn[];
    uint64_t *s_words = s_dyn, *s_pairs = s_dyn + ts_phys(CAP);
    uint32_t *s_red = reinterpret_cast<uint32_t *>(s_pairs + ts_phys(CAP / 2));
    if (n_dev && __ldg(n_dev) > capacity) return;
    for (int i = blockIdx.x; i < tiles; i += gridDim.x) {
        const int tile = (int)__ldg(tile_order + i);
        const uint2 range = ranges[tile];
        const uint32_t n = range.y - range.x;
        if (n < n_lo) break;                            // tile_order is sorted by length class: nothing longer follows
        if (n >= n_hi) continue;                        // the long class's
        __syncthreads();
        tile_sort_one_q32<THREADS, CAP>(tile, range, seg, keys_out, vals_out, s_words, s_pairs, s_red);
    }
}
constexpr size_t ts_mid_smem_bytes(int threads, int cap) {
    return (size_t)(ts_phys(cap) + ts_phys(cap / 2)) * sizeof(uint64_t) + (2 * (threads / 32) + 1) * sizeof(uint32_t);
}

template <int THREADS, int CAP>
__global__ void __launch_bounds__(THREADS) tile_sort_long_kernel(uint32_t n_lo, int tiles, uint32_t capacity, const uint32_t *__restrict__ n_dev,
                                                                 const uint2 *__restrict__ ranges, const uint32_t *__restrict__ tile_order,
                                                                 uint64_t *seg, uint64_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out) {
    extern __shared__ uint64_t s_dyn[];
    if (n_dev && __ldg(n_dev) > capacity) return;
    for (int i = blockIdx.x; i < tiles; i += gridDim.x) {
        const int tile = (int)__ldg(tile_order + i);
        const uint2 range = ranges[tile];
        if (range.y - range.x < n_lo) break;            // tile_order is sorted by length class: nothing longer follows
        __syncthreads();
        tile_sort_one<THREADS, CAP, true>(tile, range, seg, keys_out, vals_out, s_dyn);
    }
}

#ifndef LVDGS_TS_SHORT_THREADS
#define LVDGS_TS_SHORT_THREADS 256
#endif
constexpr int TS_SHORT_THREADS = LVDGS_TS_SHORT_THREADS, TS_SHORT_CAP = 2048;    // lists of 1 .. 2047 instances
constexpr int TS_MID_THREADS = 512, TS_MID_CAP = 8192;                            // lists of 2048 .. 8191
constexpr int TS_LONG_THREADS = 1024, TS_LONG_CAP = 16384;                        // 8192 and longer (beyond 16384: chunked)

// long_lists: launch the middle and long classes (lists of TS_SHORT_CAP or more).  Without them such lists are NOT sorted:
// the caller must check the longest list (binning_prep leaves it next to R) and re-run with long_lists = true.
int tile_sort_long_threshold() { return TS_SHORT_CAP; }
int launch_tile_sort(int tiles, int64_t capacity, const uint32_t *n_dev, const uint2 *ranges, const uint32_t *tile_order,
                     uint64_t *seg, uint64_t *keys_out, uint32_t *vals_out, bool long_lists, cudaStream_t s) {
    if (tiles <= 0) return 0;
    const uint32_t cap = (uint32_t)min(capacity, (int64_t)0xffffffffll);
    static std::atomic<int> sm_counts[MAX_DEVICES];          // per device (see common.cuh): attributes set + SM count
    const int dev_id = current_device();
    int sm_count = sm_counts[dev_id].load(std::memory_order_acquire);
    auto long_k = tile_sort_long_kernel<TS_LONG_THREADS, TS_LONG_CAP>;
    auto mid_k = tile_sort_mid_kernel<TS_MID_THREADS, TS_MID_CAP>;
    constexpr size_t mid_smem = ts_mid_smem_bytes(TS_MID_THREADS, TS_MID_CAP);
    if (!sm_count) {
        LVDGS_CHECK(cudaFuncSetAttribute(long_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ts_smem_bytes(TS_LONG_CAP)));
        LVDGS_CHECK(cudaFuncSetAttribute(mid_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mid_smem));
        LVDGS_CHECK(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev_id));
        sm_counts[dev_id].store(sm_count, std::memory_order_release);
    }
    if (long_lists) {
        LVDGS_PRE(s);
        long_k<<<min(tiles, sm_count), TS_LONG_THREADS, ts_smem_bytes(TS_LONG_CAP), s>>>((uint32_t)TS_MID_CAP, tiles, cap, n_dev, ranges, tile_order, seg, keys_out, vals_out);
        LVDGS_LAUNCHED(s, "tile_sort_long");
        LVDGS_PRE(s);
        mid_k<<<min(tiles, 2 * sm_count), TS_MID_THREADS, mid_smem, s>>>((uint32_t)TS_SHORT_CAP, (uint32_t)TS_MID_CAP, tiles, cap, n_dev, ranges, tile_order, seg, keys_out, vals_out);
        LVDGS_LAUNCHED(s, "tile_sort_mid");
    }
    LVDGS_PRE(s);
    LVDGS_CHECK(launch_after_kernel(tile_sort_short_kernel<TS_SHORT_THREADS, TS_SHORT_CAP>, dim3(tiles), dim3(TS_SHORT_THREADS), 0, s, (uint32_t)TS_SHORT_CAP,
                                    long_lists ? 0 : 1, cap, n_dev, ranges, tile_order, seg, keys_out, vals_out));
    LVDGS_LAUNCHED(s, "tile_sort");
    return 0;
}

}  // namespace lvdgs
