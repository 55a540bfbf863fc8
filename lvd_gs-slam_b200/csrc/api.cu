// api.cu -- the extern "C" entry points declared in include/lvdgs.h: buffer layouts, argument checks and the
// launch sequence of the forward and backward passes.  No torch types, no allocation, one host sync (R).
#include "common.cuh"
#include <cstdlib>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace lvdgs {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};
thread_local int g_debug_sync = 0;
static thread_local int64_t t_tail_reruns = 0;     // forwards of this thread whose speculative tail had to be repeated

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// ---- optional per-launch profiler: one CUDA event after every launch, on the launching stream ----
// The profiler belongs to the thread that switched it on: other threads (another engine, another device) launch unprofiled.
struct ProfEntry { const char *name; cudaEvent_t ev; cudaEvent_t pre; };
static thread_local std::vector<ProfEntry> g_prof;
static thread_local bool g_prof_on = false;
static thread_local cudaEvent_t g_prof_pending_pre = nullptr;

static bool pdl_profiling();
bool pdl_enabled() {
    static const bool on = [] { const char *e = getenv("LVDGS_PDL"); return !(e && e[0] == '0'); }();
    return on && !g_prof_on;           // the profiler's events between launches serialise anyway
}

static bool pdl_profiling() { return g_prof_on; }

int profile_pre(cudaStream_t s) {
    if (!g_prof_on) return 0;
    if (!g_prof_pending_pre) LVDGS_CHECK(cudaEventCreate(&g_prof_pending_pre));
    LVDGS_CHECK(cudaEventRecord(g_prof_pending_pre, s));
    return 0;
}

int profile_mark(const char *name, cudaStream_t s) {
    if (!g_prof_on) return 0;
    ProfEntry e{name, nullptr, g_prof_pending_pre};
    g_prof_pending_pre = nullptr;
    LVDGS_CHECK(cudaEventCreate(&e.ev));
    LVDGS_CHECK(cudaEventRecord(e.ev, s));
    g_prof.push_back(e);
    return 0;
}

static void geom_layout(int32_t P, lvdgs_geom_layout &l) {
    size_t o = 0;
    const size_t n = (size_t)(P > 0 ? P : 1);
    l.depths = o; o += align_up(n * sizeof(float));
    l.means2D = o; o += align_up(n * sizeof(float4));
    l.conic_opacity = o; o += align_up(n * sizeof(float4));
    l.rgbd = o; o += align_up(n * sizeof(float4));
    l.rect = o; o += align_up(n * sizeof(short4));
    l.tiles_touched = o; o += align_up(n * sizeof(uint32_t));
    l.point_offsets = o; o += align_up(n * sizeof(uint32_t));
    l.clamped = o; o += align_up(n * sizeof(uint8_t));
    l.scan_state = o; o += align_up(((n + 255) / 256) * sizeof(uint32_t)) + 256;   // block sums + counters
    l.visible_list = o; o += align_up(n * sizeof(uint32_t));
    l.total = o;
}
static void binning_layout(int64_t R, lvdgs_binning_layout &l) {
    size_t o = 0;
    const size_t n = (size_t)(R > 0 ? R : 1);
    for (int k = 0; k < 2; ++k) { l.keys[k] = o; o += align_up(n * sizeof(uint64_t)); }
    for (int k = 0; k < 2; ++k) { l.vals[k] = o; o += align_up(n * sizeof(uint32_t)); }
    l.sort_ws = o; o += align_up(sort_workspace_bytes((int64_t)n));
    l.sorted_sel = o; o += 256;
    l.total = o;
}
static void img_layout(int32_t W, int32_t H, lvdgs_img_layout &l) {
    size_t o = 0;
    const size_t n = (size_t)W * (size_t)H;
    const size_t tiles = (size_t)((W + TILE - 1) / TILE) * (size_t)((H + TILE - 1) / TILE);
    l.final_T = o; o += align_up(n * sizeof(float));
    l.n_contrib = o; o += align_up(n * sizeof(uint32_t));
    l.ranges = o; o += align_up(tiles * sizeof(uint2));
    const size_t gridn = (size_t)((W + TILE - 1) / TILE + 1) * (size_t)((H + TILE - 1) / TILE + 1);
    l.tile_order = o; o += align_up(tiles * sizeof(uint32_t));
    l.tile_grid = o; o += align_up(gridn * sizeof(int32_t));
    l.sort_hist = o; o += align_up(SORT_MAX_PASSES * SORT_BINS * sizeof(uint32_t));
    l.tile_cursor = o; o += align_up(tiles * CURSOR_STRIDE * sizeof(uint32_t));
    l.total = o;
}

static GeomPtrs geom_ptrs(void *base, int32_t P) {
    lvdgs_geom_layout l; geom_layout(P, l);
    char *b = (char *)base;
    GeomPtrs g;
    g.depths = (float *)(b + l.depths); g.means2D = (float4 *)(b + l.means2D);
    g.conic_opacity = (float4 *)(b + l.conic_opacity); g.rgbd = (float4 *)(b + l.rgbd);
    g.rect = (short4 *)(b + l.rect); g.tiles_touched = (uint32_t *)(b + l.tiles_touched);
    g.point_offsets = (uint32_t *)(b + l.point_offsets); g.clamped = (uint8_t *)(b + l.clamped);
    g.block_sums = (uint32_t *)(b + l.scan_state);
    g.num_instances = (uint32_t *)(b + l.visible_list - 256);
    g.visible_list = (uint32_t *)(b + l.visible_list);
    return g;
}
static BinPtrs bin_ptrs(void *base, int64_t R) {
    lvdgs_binning_layout l; binning_layout(R, l);
    char *b = (char *)base;
    BinPtrs p;
    for (int k = 0; k < 2; ++k) { p.keys[k] = (uint64_t *)(b + l.keys[k]); p.vals[k] = (uint32_t *)(b + l.vals[k]); }
    p.sort_ws = b + l.sort_ws; p.sorted_sel = (int32_t *)(b + l.sorted_sel);
    return p;
}
static ImgPtrs img_ptrs(void *base, int32_t W, int32_t H) {
    lvdgs_img_layout l; img_layout(W, H, l);
    char *b = (char *)base;
    ImgPtrs p;
    p.final_T = (float *)(b + l.final_T); p.n_contrib = (uint32_t *)(b + l.n_contrib); p.ranges = (uint2 *)(b + l.ranges);
    p.tile_order = (uint32_t *)(b + l.tile_order); p.tile_grid = (int32_t *)(b + l.tile_grid); p.sort_hist = (uint32_t *)(b + l.sort_hist);
    p.tile_cursor = (uint32_t *)(b + l.tile_cursor);
    return p;
}

static int check_params(const lvdgs_raster_params *p) {
    if (!p) { set_error("params is NULL"); return 1; }
    if (p->P < 0 || p->width <= 0 || p->height <= 0) { set_error("bad sizes P=%d W=%d H=%d", p->P, p->width, p->height); return 1; }
    if (p->sh_degree < 0 || p->sh_degree > 3) { set_error("sh_degree %d not in 0..3", p->sh_degree); return 1; }
    if (!(p->tan_fovx > 0.f) || !(p->tan_fovy > 0.f)) { set_error("tan_fov must be positive"); return 1; }
    const int64_t tiles = (int64_t)((p->width + TILE - 1) / TILE) * ((p->height + TILE - 1) / TILE);
    if ((p->width + TILE - 1) / TILE > 32767 || (p->height + TILE - 1) / TILE > 32767 || tiles > (1ll << 31)) {
        set_error("image too large for the tile grid"); return 1;
    }
    return 0;
}

}  // namespace lvdgs

using namespace lvdgs;

extern "C" {

int lvdgs_version(void) { return 100; }
const char *lvdgs_last_error(void) { return g_err; }
int lvdgs_set_device(int device) { LVDGS_CHECK(cudaSetDevice(device)); return 0; }
int64_t lvdgs_launch_count(void) { return g_launches.load(); }
int64_t lvdgs_tail_rerun_count(void) { return t_tail_reruns; }
void lvdgs_reset_launch_count(void) { g_launches.store(0); }

int lvdgs_profile_begin(void *stream) {
    for (auto &e : g_prof) { cudaEventDestroy(e.ev); if (e.pre) cudaEventDestroy(e.pre); }
    g_prof.clear();
    g_prof_on = true;
    return profile_mark("(begin)", (cudaStream_t)stream);
}

int lvdgs_profile_end(void *stream, char *names, size_t names_bytes, float *ms, int32_t max_entries) {
    g_prof_on = false;
    LVDGS_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
    std::string all;
    int n = 0;
    for (size_t i = 1; i < g_prof.size() && n < max_entries; ++i, ++n) {
        float t = 0.f;
        // a launch has its own "pre" event; host-side markers are measured from the previous entry
        LVDGS_CHECK(cudaEventElapsedTime(&t, g_prof[i].pre ? g_prof[i].pre : g_prof[i - 1].ev, g_prof[i].ev));
        ms[n] = t;
        all += g_prof[i].name;
        all += '\n';
    }
    if (names && names_bytes) {
        const size_t c = all.size() < names_bytes - 1 ? all.size() : names_bytes - 1;
        memcpy(names, all.data(), c);
        names[c] = 0;
    }
    for (auto &e : g_prof) { cudaEventDestroy(e.ev); if (e.pre) cudaEventDestroy(e.pre); }
    g_prof.clear();
    return n;
}

int lvdgs_get_geom_layout(int32_t P, lvdgs_geom_layout *out) { if (!out) return 1; geom_layout(P, *out); return 0; }
int lvdgs_get_binning_layout(int64_t R, lvdgs_binning_layout *out) { if (!out) return 1; binning_layout(R, *out); return 0; }
int lvdgs_get_img_layout(int32_t W, int32_t H, lvdgs_img_layout *out) { if (!out) return 1; img_layout(W, H, *out); return 0; }

// pinned slot + event for the asynchronous read-back of the instance count: one per host thread AND device (an event
// belongs to the device it was created on)
struct ReadBack { uint32_t *pinned = nullptr; cudaEvent_t event = nullptr; uint32_t seq = 0; };
// LVDGS_POLL_R=0: read (R, longest list) back with a copy + event instead of the kernel's direct write into pinned memory
static bool poll_readback() {
    static const bool on = [] { const char *e = getenv("LVDGS_POLL_R"); return !(e && e[0] == '0'); }();
    return on;
}
// waits until binning_prep has published sequence number `seq` in the pinned slot (it wrote R and the longest list first)
static int wait_readback(const uint32_t *pinned, uint32_t seq, cudaStream_t s) {
    const volatile uint32_t *flag = pinned + 2;
    for (uint64_t spins = 1; *flag != seq; ++spins) {
        if ((spins & 0xfff) == 0) {              // every few thousand polls: has the stream died or drained without publishing?
            const cudaError_t q = cudaStreamQuery(s);
            if (q == cudaSuccess) { if (*flag == seq) break; set_error("forward: the instance count never arrived from the device"); return 1; }
            if (q != cudaErrorNotReady) { set_error("forward: %s while waiting for the instance count", cudaGetErrorString(q)); return 1; }
        }
    }
    std::atomic_thread_fence(std::memory_order_acquire);      // R and the longest list were written before the sequence number
    return 0;
}
static thread_local ReadBack t_readback[MAX_DEVICES];
// longest tile lists of this thread's last forwards (read back with R), per (device, image size): decide whether the next
// forward launches tile_sort's long-list classes speculatively (a wrong guess costs time, never correctness).  The
// maximum over the last eight forwards OF THE SAME SHAPE is used, so cameras with and without long lists rendered in turn
// (a mapping window), or a full-resolution tracking render alternating with the 512x144 depth render of
// utils/init_pose.py:145, never re-run each other's tails.
struct LongestHist { uint64_t key = ~0ull; uint32_t v[8] = {0xffffffffu, 0, 0, 0, 0, 0, 0, 0}; int pos = 0; };
static thread_local LongestHist t_longest[8];      // a handful of (device, W, H) combinations per thread, LRU by slot 0
static LongestHist &longest_for(int dev, int W, int H) {
    const uint64_t key = ((uint64_t)dev << 48) | ((uint64_t)(uint32_t)W << 24) | (uint64_t)(uint32_t)H;
    for (int i = 0; i < 8; ++i)
        if (t_longest[i].key == key) {
            if (i) { LongestHist h = t_longest[i]; for (int j = i; j > 0; --j) t_longest[j] = t_longest[j - 1]; t_longest[0] = h; }
            return t_longest[0];
        }
    for (int j = 7; j > 0; --j) t_longest[j] = t_longest[j - 1];
    t_longest[0] = LongestHist();
    t_longest[0].key = key;
    return t_longest[0];
}
static uint32_t longest_recent(const LongestHist &h) {
    uint32_t m = 0;
    for (uint32_t v : h.v) m = v > m ? v : m;
    return m;
}

// everything after the instance count is known on the DEVICE: keys, sort, ranges, blend.  `capacity` sizes the
// launches and the binning arena; the kernels clamp to min(R, capacity) read from device memory.
static int launch_bin_and_blend(const lvdgs_raster_params &p, const GeomPtrs &g, const BinPtrs &b, const ImgPtrs &im,
                                int64_t capacity, bool rerun, bool long_lists, const float *background, float *out_color, float *out_depth,
                                float *out_opacity, int32_t *n_touched, cudaStream_t s) {
    const int W = p.width, H = p.height;
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    const uint32_t *R_dev = g.num_instances;
    // n_touched was zeroed by preprocess_forward; a re-run after a failed speculative launch counts again
    if (rerun) LVDGS_CHECK(cudaMemsetAsync(n_touched, 0, sizeof(int32_t) * (size_t)p.P, s));
    int sel = 0;
    // a re-run after a failed speculative launch: the emission counts the visible list (and claims tile slots) again
    if (rerun) LVDGS_CHECK(cudaMemsetAsync(g.num_instances + 2, 0, sizeof(uint32_t), s));
    if (p.flags & LVDGS_FLAG_GLOBAL_SORT) {
        if (launch_emit_keys(p.P, W, H, g, capacity, b.keys[0], b.vals[0], nullptr, nullptr, nullptr, s)) return 1;
        const int end_bit = 32 + tile_bits((uint32_t)(gx * gy));
        if (launch_sort_pairs(capacity, R_dev, b.keys[0], b.keys[1], b.vals[0], b.vals[1], end_bit, b.sort_ws,
                              sort_workspace_bytes(capacity), im.sort_hist, &sel, s)) return 1;
    } else {
        // instances go straight into their tile's segment of keys[0]; one CTA per tile sorts it into keys[1] / vals[1].
        // The cursors were zeroed together with the tile grid
        if (rerun) LVDGS_CHECK(cudaMemsetAsync(im.tile_cursor, 0, sizeof(uint32_t) * CURSOR_STRIDE * (size_t)gx * gy, s));
        if (launch_emit_keys(p.P, W, H, g, capacity, b.keys[0], nullptr, im.tile_cursor, im.ranges, b.sorted_sel, s)) return 1;
        if (launch_tile_sort(gx * gy, capacity, R_dev, im.ranges, im.tile_order, b.keys[0], b.keys[1], b.vals[1], long_lists, s)) return 1;
        sel = 1;
    }
    // which of the two key / value buffers holds the sorted list (read by the tests' introspection only): the tile-segment
    // path's emission kernel wrote it; the onesweep selector is known here
    if (p.flags & LVDGS_FLAG_GLOBAL_SORT) LVDGS_CHECK(cudaMemcpyAsync(b.sorted_sel, &sel, sizeof(int32_t), cudaMemcpyHostToDevice, s));
    return launch_blend_forward(W, H, capacity, R_dev, im.ranges, b.vals[sel], g, im.tile_order, background, out_color, out_depth, out_opacity,
                                im.final_T, im.n_contrib, n_touched, s);
}

int lvdgs_rasterize_forward(const lvdgs_raster_params *prm, const float *background, const float *means3D,
                            const float *colors_precomp, const float *opacities, const float *scales,
                            const float *rotations, const float *cov3D_precomp, const float *viewmatrix,
                            const float *projmatrix, const float *projmatrix_raw, const float *shs,
                            const float *campos, lvdgs_resize_fn resize, void *resize_user, int64_t capacity_hint,
                            float *out_color, int32_t *radii, float *out_depth, float *out_opacity,
                            int32_t *n_touched, int64_t *num_rendered, int64_t *binning_capacity, void *stream) {
    (void)projmatrix_raw;
    if (check_params(prm)) return 1;
    const lvdgs_raster_params &p = *prm;
    if (p.P > 0) {
        if ((shs == nullptr) == (colors_precomp == nullptr)) { set_error("provide exactly one of shs / colors_precomp"); return 1; }
        const bool has_sr = scales != nullptr && rotations != nullptr;
        if (has_sr == (cov3D_precomp != nullptr) || ((scales == nullptr) != (rotations == nullptr))) {
            set_error("provide exactly one of (scales, rotations) / cov3D_precomp"); return 1;
        }
    }
    if (shs && p.sh_coeffs < (p.sh_degree + 1) * (p.sh_degree + 1)) { set_error("sh_coeffs %d too small for degree %d", p.sh_coeffs, p.sh_degree); return 1; }
    if (!resize || !out_color || !out_depth || !out_opacity || !num_rendered || !binning_capacity || !background ||
        !viewmatrix || !projmatrix || !campos) {
        set_error("NULL required argument"); return 1;
    }
    if (p.P > 0 && (!means3D || !opacities || !radii || !n_touched)) { set_error("NULL per-Gaussian argument"); return 1; }
    cudaStream_t s = (cudaStream_t)stream;
    g_debug_sync = p.debug;
    const int W = p.width, H = p.height;
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    *num_rendered = 0;
    *binning_capacity = 0;

    lvdgs_img_layout il; img_layout(W, H, il);
    void *img_base = resize(resize_user, LVDGS_BUF_IMG, il.total);
    if (!img_base) { set_error("resize callback returned NULL (img)"); return 1; }
    ImgPtrs im = img_ptrs(img_base, W, H);

    if (p.P == 0) {     // empty map: background only
        LVDGS_CHECK(cudaMemsetAsync(im.ranges, 0, sizeof(uint2) * (size_t)gx * gy, s));
        GeomPtrs g{};
        return launch_blend_forward(W, H, 0, nullptr, im.ranges, nullptr, g, nullptr, background, out_color, out_depth, out_opacity, im.final_T,
                                    im.n_contrib, n_touched, s);
    }
    const int dev_id = current_device();
    ReadBack &rb = t_readback[dev_id];
    if (!rb.pinned) {
        LVDGS_CHECK(cudaHostAlloc((void **)&rb.pinned, 64, cudaHostAllocDefault));      // UVA: the device can write it directly
        memset(rb.pinned, 0, 64);
        LVDGS_CHECK(cudaEventCreateWithFlags(&rb.event, cudaEventDisableTiming));
    }
    uint32_t *const t_pinned_R = rb.pinned;
    const cudaEvent_t t_R_event = rb.event;
    LongestHist &lh = longest_for(dev_id, W, H);
    lvdgs_geom_layout gl; geom_layout(p.P, gl);
    void *geom_base = resize(resize_user, LVDGS_BUF_GEOM, gl.total);
    if (!geom_base) { set_error("resize callback returned NULL (geom)"); return 1; }
    GeomPtrs g = geom_ptrs(geom_base, p.P);
    // the tile grid + digit histograms + cursors are accumulated with atomics: preprocess_forward clears them (no memset launch)
    if (launch_preprocess_forward(p, means3D, colors_precomp, opacities, scales, rotations, cov3D_precomp, viewmatrix,
                                  projmatrix, shs, campos, radii, n_touched, g, im, s)) return 1;
    // block offsets, R, tile ranges and all digit histograms of the sort, from the block sums and per-tile counts
    // (R, longest list) reach the host either by the kernel's own write into pinned memory + a sequence number the host
    // polls (default: nothing sits between binning_prep and the key emission), or by a copy + event
    const bool poll = poll_readback() && !pdl_profiling();
    const uint32_t seq = poll ? ++rb.seq : 0u;
    if (launch_binning_prep(p.P, W, H, (p.flags & LVDGS_FLAG_GLOBAL_SORT) ? 32 + tile_bits((uint32_t)(gx * gy)) : 0, g, im,
                            poll ? t_pinned_R : nullptr, seq, s)) return 1;
    if (!poll) {
        LVDGS_CHECK(cudaMemcpyAsync(t_pinned_R, g.num_instances, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        LVDGS_CHECK(cudaEventRecord(t_R_event, s));
    }

    // Speculative launch: with a capacity hint the whole rest of the forward is queued BEFORE the host waits for R,
    // so the device never idles across the read-back.  If R turns out larger than the hint, the tail is re-run with
    // an exactly sized arena (the speculative results are simply overwritten).
    int64_t capacity = capacity_hint > 0 ? capacity_hint : 0;
    bool launched = false, launched_long = false;
    if (capacity > 0) {
        lvdgs_binning_layout bl; binning_layout(capacity, bl);
        void *bin_base = resize(resize_user, LVDGS_BUF_BINNING, bl.total);
        if (!bin_base) { set_error("resize callback returned NULL (binning)"); return 1; }
        // guess from the previous forward, with hysteresis
        launched_long = (p.flags & LVDGS_FLAG_GLOBAL_SORT) || longest_recent(lh) >= (uint32_t)(tile_sort_long_threshold() * 3 / 4);
        if (launch_bin_and_blend(p, g, bin_ptrs(bin_base, capacity), im, capacity, false, launched_long, background, out_color, out_depth,
                                 out_opacity, n_touched, s)) return 1;
        launched = true;
    }
    if (poll) { if (wait_readback(t_pinned_R, seq, s)) return 1; }
    else LVDGS_CHECK(cudaEventSynchronize(t_R_event));
    if (profile_mark("(host: R read-back)", s)) return 1;
    const int64_t R = t_pinned_R[0];
    const uint32_t longest = t_pinned_R[1];
    lh.pos = (lh.pos + 1) & 7;
    lh.v[lh.pos] = longest;
    *num_rendered = R;
    const bool need_long = longest >= (uint32_t)tile_sort_long_threshold();      // exact: just read back
    if (!launched || R > capacity || (need_long && !launched_long)) {
        if (launched) ++t_tail_reruns;
        if (!launched || R > capacity) capacity = R > 0 ? R : 1;
        lvdgs_binning_layout bl; binning_layout(capacity, bl);
        void *bin_base = resize(resize_user, LVDGS_BUF_BINNING, bl.total);
        if (!bin_base) { set_error("resize callback returned NULL (binning)"); return 1; }
        if (launch_bin_and_blend(p, g, bin_ptrs(bin_base, capacity), im, capacity, launched, need_long, background, out_color, out_depth,
                                 out_opacity, n_touched, s)) return 1;
    }
    *binning_capacity = capacity;
    return 0;
}

void *lvdgs_static_resize(void *user, int32_t which, size_t bytes) {
    lvdgs_static_buffers *b = (lvdgs_static_buffers *)user;
    if (!b || which < 0 || which > 2) return nullptr;
    if (b->base[which] && bytes <= b->capacity[which]) return b->base[which];
    return b->fallback ? b->fallback(b->fallback_user, which, bytes) : nullptr;
}

int lvdgs_zero_async(void *ptr, size_t bytes, void *stream) {
    if (!ptr && bytes) { set_error("zero_async: NULL pointer"); return 1; }
    if (bytes) LVDGS_CHECK(cudaMemsetAsync(ptr, 0, bytes, (cudaStream_t)stream));
    return 0;
}

size_t lvdgs_backward_scratch_bytes(int32_t P, int64_t R) {
    (void)R;
    return align_up((size_t)(P > 0 ? P : 1) * ACC_STRIDE * sizeof(float));
}

int lvdgs_rasterize_backward(const lvdgs_raster_params *prm, const float *background, const float *means3D,
                             const int32_t *radii, const float *colors_precomp, const float *opacities,
                             const float *scales, const float *rotations, const float *cov3D_precomp,
                             const float *viewmatrix, const float *projmatrix, const float *projmatrix_raw,
                             const float *dL_dout_color, const float *dL_dout_depth, const float *dL_dout_opacity,
                             const float *shs, const float *campos, const void *geom_buffer, int64_t R,
                             int64_t binning_capacity, const void *binning_buffer, const void *img_buffer, void *scratch,
                             size_t scratch_bytes, float *dL_dmeans2D, float *dL_dcolors, float *dL_dopacity,
                             float *dL_dmeans3D, float *dL_dcov3D, float *dL_dsh, float *dL_dscales,
                             float *dL_drots, float *dL_dtau, float *dL_dtau_sum, void *stream) {
    (void)opacities;
    if (check_params(prm)) return 1;
    const lvdgs_raster_params &p = *prm;
    cudaStream_t s = (cudaStream_t)stream;
    g_debug_sync = p.debug;
    if (p.P == 0) {
        if (dL_dtau_sum) LVDGS_CHECK(cudaMemsetAsync(dL_dtau_sum, 0, 6 * sizeof(float), s));
        return 0;
    }
    if (!geom_buffer || !img_buffer || (R > 0 && !binning_buffer) || !scratch || !dL_dout_color) { set_error("NULL buffer"); return 1; }
    if (scratch_bytes < lvdgs_backward_scratch_bytes(p.P, R)) { set_error("backward scratch too small"); return 1; }
    if (!(p.flags & LVDGS_FLAG_POSE_ONLY)) {
        if (!dL_dopacity || !dL_dmeans3D) { set_error("NULL gradient output"); return 1; }
        if (colors_precomp && !dL_dcolors) { set_error("dL_dcolors required with colors_precomp"); return 1; }
        if (cov3D_precomp && !dL_dcov3D) { set_error("dL_dcov3D required with cov3D_precomp"); return 1; }
        if (!colors_precomp && !dL_dsh) { set_error("dL_dsh required when shs are used"); return 1; }
        if (!cov3D_precomp && (!dL_dscales || !dL_drots)) { set_error("dL_dscales / dL_drots required"); return 1; }
    }
    const int W = p.width, H = p.height;
    GeomPtrs g = geom_ptrs(const_cast<void *>(geom_buffer), p.P);
    ImgPtrs im = img_ptrs(const_cast<void *>(img_buffer), W, H);
    BlendGradPtrs bg{(float *)scratch};
    if (!(p.flags & LVDGS_FLAG_ZEROED_SCRATCH)) LVDGS_CHECK(cudaMemsetAsync(scratch, 0, (size_t)p.P * ACC_STRIDE * sizeof(float), s));
    // dL_dtau_sum is zeroed by the blend backward (first thread of the launch) when there is one
    bool tau_zeroed = false;
    if (R > 0) {
        BinPtrs b = bin_ptrs(const_cast<void *>(binning_buffer), binning_capacity);
        const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
        const int passes = (32 + tile_bits((uint32_t)(gx * gy)) + 7) / 8;
        const int sel = (p.flags & LVDGS_FLAG_GLOBAL_SORT) ? (passes & 1) : 1;
        // tracking: the pose gradient needs neither the colour sums (colours do not depend on the pose at SH degree 0 /
        // with precomputed colours) nor, without a depth loss, the depth sum -> the six geometric moments only
        const bool moments_only = (p.flags & LVDGS_FLAG_POSE_ONLY) && !dL_dout_depth && (p.sh_degree == 0 || colors_precomp);
        if (launch_blend_backward(p.P, W, H, R, im.ranges, b.vals[sel], im.tile_order, g, background, im.final_T, im.n_contrib,
                                  dL_dout_color, dL_dout_depth, dL_dout_opacity, p.flags, moments_only, bg, dL_dtau_sum, s)) return 1;
        tau_zeroed = true;
    }
    if (launch_preprocess_backward(p, means3D, radii, shs, scales, rotations, cov3D_precomp, viewmatrix, projmatrix,
                                   projmatrix_raw, campos, g, bg, colors_precomp != nullptr, dL_dmeans2D, dL_dcolors,
                                   dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drots, dL_dtau,
                                   dL_dtau_sum, tau_zeroed, s)) return 1;
    return 0;
}

int lvdgs_mark_visible(int32_t P, const float *means3D, const float *viewmatrix, const float *projmatrix,
                       uint8_t *present, void *stream) {
    (void)projmatrix;
    if (P < 0 || (P > 0 && (!means3D || !viewmatrix || !present))) { set_error("mark_visible: bad arguments"); return 1; }
    g_debug_sync = 0;
    return launch_mark_visible(P, means3D, viewmatrix, present, (cudaStream_t)stream);
}

size_t lvdgs_dist2_workspace_bytes(int32_t P) { return dist2_workspace_bytes(P); }
int lvdgs_dist2(int32_t P, const float *points, float *mean_dists, void *workspace, size_t workspace_bytes,
                void *stream) {
    if (P < 0 || (P > 0 && (!points || !mean_dists))) { set_error("dist2: bad arguments"); return 1; }
    g_debug_sync = 0;
    return launch_dist2(P, points, mean_dists, workspace, workspace_bytes, (cudaStream_t)stream);
}

int lvdgs_adam_step(int64_t n, float *params, const float *grads, float *exp_avg, float *exp_avg_sq, int32_t groups,
                    const int64_t *group_end, const float *lr, double beta1, double beta2, double eps, int32_t step,
                    void *stream) {
    if (!params || !grads || !exp_avg || !exp_avg_sq || !group_end || !lr || step < 1) { set_error("adam: bad arguments"); return 1; }
    g_debug_sync = 0;
    return launch_adam_step(n, params, grads, exp_avg, exp_avg_sq, groups, group_end, lr, beta1, beta2, eps, step,
                            (cudaStream_t)stream);
}

int lvdgs_exchange_adam(int32_t world, int32_t rank, const float *const *grad_ptrs, float *const *param_ptrs, float *const *act_ptrs,
                        int64_t lo, int64_t hi, float *exp_avg, float *exp_avg_sq, int32_t groups, const int64_t *group_end,
                        const float *lr, const int64_t *act_offsets, int64_t act_total, double beta1, double beta2, double eps,
                        int32_t step, const float *mc_grad, float *mc_param, float *mc_act, int32_t act_mode, void *stream) {
    if (!grad_ptrs || !param_ptrs || !exp_avg || !exp_avg_sq || !group_end || !lr || step < 1) { set_error("exchange: bad arguments"); return 1; }
    g_debug_sync = 0;
    return launch_exchange_adam(world, rank, grad_ptrs, param_ptrs, act_ptrs, lo, hi, exp_avg, exp_avg_sq, groups, group_end, lr,
                                act_offsets, act_total, beta1, beta2, eps, step, mc_grad, mc_param, mc_act, act_mode, (cudaStream_t)stream);
}

size_t lvdgs_fused_loss_workspace_bytes(void) { return fused_loss_workspace_bytes(); }
int lvdgs_fused_loss(int32_t width, int32_t height, const float *color, const float *depth, const float *opacity,
                     const float *gt_color, const float *gt_depth, const float *grad_mask, const float *exposure,
                     float rgb_boundary_threshold, float w_rgb, float w_depth, int32_t flags, float *g_color,
                     float *g_depth, float *g_opacity, float *out, void *workspace, size_t workspace_bytes, void *stream) {
    if (width <= 0 || height <= 0 || !color || !gt_color || !g_color || !out || !workspace) { set_error("fused_loss: bad arguments"); return 1; }
    if (w_depth != 0.f && gt_depth && !depth) { set_error("fused_loss: depth term without a rendered depth"); return 1; }
    if ((flags & (LVDGS_LOSS_OPACITY_WEIGHT | LVDGS_LOSS_DEPTH_NEEDS_OPAQUE)) && !opacity) { set_error("fused_loss: flags need the opacity image"); return 1; }
    g_debug_sync = 0;
    return launch_fused_loss(width, height, color, depth, opacity, gt_color, gt_depth, grad_mask, exposure,
                             rgb_boundary_threshold, w_rgb, w_depth, flags, g_color, g_depth, g_opacity, out, workspace,
                             workspace_bytes, (cudaStream_t)stream);
}

size_t lvdgs_masked_ssim_loss_workspace_bytes(int32_t width, int32_t height) { return masked_ssim_workspace_bytes(width, height); }
int lvdgs_masked_ssim_loss(int32_t width, int32_t height, const float *image, const float *gt_image, const uint8_t *static_mask,
                           const float *background, const float *depth, const float *mono_depth, float lambda_dssim,
                           float depth_lambda, float *g_image, float *g_depth, float *out, void *workspace, size_t workspace_bytes,
                           void *stream) {
    if (width <= 0 || height <= 0 || !image || !gt_image || !background || !g_image || !out || !workspace) { set_error("masked_ssim_loss: bad arguments"); return 1; }
    g_debug_sync = 0;
    return launch_masked_ssim_loss(width, height, image, gt_image, static_mask, background, depth, mono_depth, lambda_dssim,
                                   depth_lambda, g_image, g_depth, out, workspace, workspace_bytes, (cudaStream_t)stream);
}

int lvdgs_covis_counts(int64_t n, const void *a, const void *b, int32_t elem_bytes, uint64_t *out, void *stream) {
    if (n < 0 || !out || (n > 0 && (!a || !b))) { set_error("covis_counts: bad arguments"); return 1; }
    g_debug_sync = 0;
    return launch_covis(n, a, b, elem_bytes, (unsigned long long *)out, (cudaStream_t)stream);
}
int lvdgs_n_obs(int64_t n, int32_t K, const void *const *masks, int32_t elem_bytes, int32_t *n_obs, void *stream) {
    if (n < 0 || K < 0 || (n > 0 && (!n_obs || (K > 0 && !masks)))) { set_error("n_obs: bad arguments"); return 1; }
    g_debug_sync = 0;
    return launch_n_obs(n, K, masks, elem_bytes, n_obs, (cudaStream_t)stream);
}

size_t lvdgs_compact_workspace_bytes(int64_t n) { return compact_workspace_bytes(n > 0 ? n : 0); }
int lvdgs_compact_count(int64_t n, const uint8_t *keep, void *workspace, size_t workspace_bytes, uint32_t **count_dev,
                        void *stream) {
    if (n < 0 || !workspace || !count_dev || (n > 0 && !keep)) { set_error("compact_count: bad arguments"); return 1; }
    g_debug_sync = 0;
    return launch_compact_count(n, keep, workspace, workspace_bytes, count_dev, (cudaStream_t)stream);
}
int lvdgs_compact_move(int64_t n, const uint8_t *keep, const void *workspace, int32_t n_arrays,
                       const float *const *src, float *const *dst, const int32_t *widths, void *stream) {
    if (n < 0 || !workspace || (n_arrays > 0 && (!src || !dst || !widths)) || (n > 0 && !keep)) { set_error("compact_move: bad arguments"); return 1; }
    g_debug_sync = 0;
    return launch_compact_move(n, keep, workspace, n_arrays, src, dst, widths, (cudaStream_t)stream);
}

int lvdgs_gather_rows(int64_t n_idx, const int64_t *idx, int64_t n_src_rows, int32_t n_arrays, const float *const *src,
                      float *const *dst, const int32_t *widths, void *stream) {
    if (n_idx < 0 || n_src_rows < 0 || (n_idx > 0 && (!idx || (n_arrays > 0 && (!src || !dst || !widths))))) { set_error("gather_rows: bad arguments"); return 1; }
    g_debug_sync = 0;
    return launch_gather_rows(n_idx, idx, n_src_rows, n_arrays, src, dst, widths, (cudaStream_t)stream);
}

int lvdgs_gaussian_activate(int64_t P, const float *raw_opacity, const float *raw_scales, const float *raw_rotations, float *opacity,
                            float *scales, float *rotations, void *stream) {
    if (P < 0 || (P > 0 && (!raw_opacity || !raw_scales || !raw_rotations || !opacity || !scales || !rotations))) { set_error("gaussian_activate: bad arguments"); return 1; }
    if (((uintptr_t)raw_rotations | (uintptr_t)rotations) & 15) { set_error("gaussian_activate: rotations must be 16-byte aligned"); return 1; }
    g_debug_sync = 0;
    return launch_gaussian_activate(P, raw_opacity, raw_scales, raw_rotations, opacity, scales, rotations, (cudaStream_t)stream);
}
int lvdgs_gaussian_activation_backward(int64_t P, const float *opacity, const float *scales, const float *rotations,
                                       const float *raw_rotations, float *g_opacity, float *g_scales, float *g_rotations, void *stream) {
    if (P < 0 || (P > 0 && (!opacity || !scales || !rotations || !raw_rotations || !g_opacity || !g_scales || !g_rotations))) { set_error("gaussian_activation_backward: bad arguments"); return 1; }
    if (((uintptr_t)raw_rotations | (uintptr_t)rotations | (uintptr_t)g_rotations) & 15) { set_error("gaussian_activation_backward: rotations must be 16-byte aligned"); return 1; }
    g_debug_sync = 0;
    return launch_gaussian_activation_backward(P, opacity, scales, rotations, raw_rotations, g_opacity, g_scales, g_rotations, (cudaStream_t)stream);
}

int lvdgs_pose_step(lvdgs_pose_state *state, const float *g_tau, const float *g_exposure, float lr_rot, float lr_trans,
                    float lr_exposure, double beta1, double beta2, double eps, int32_t step, float converged_threshold,
                    void *stream) {
    if (!state || !g_tau || step < 1) { set_error("pose_step: bad arguments"); return 1; }
    g_debug_sync = 0;
    return launch_pose_step(state, g_tau, g_exposure, lr_rot, lr_trans, lr_exposure, beta1, beta2, eps, step,
                            converged_threshold, (cudaStream_t)stream);
}

int lvdgs_fp32_peak(int32_t blocks, int32_t iters, int32_t mode, float *out, double *fmas, void *stream) {
    if (blocks <= 0 || iters <= 0 || !out) { set_error("fp32_peak: bad arguments"); return 1; }
    g_debug_sync = 0;
    return launch_fp32_peak(blocks, iters, mode, out, fmas, (cudaStream_t)stream);
}

size_t lvdgs_sort_workspace_bytes(int64_t n) { return sort_workspace_bytes(n > 0 ? n : 1); }
int lvdgs_sort_pairs(int64_t n, uint64_t *keys0, uint64_t *keys1, uint32_t *vals0, uint32_t *vals1,
                     int32_t end_bit, void *workspace, size_t workspace_bytes, int32_t *selector, void *stream) {
    g_debug_sync = 0;
    int sel = 0;
    const int rc = launch_sort_pairs(n, nullptr, keys0, keys1, vals0, vals1, end_bit, workspace, workspace_bytes, nullptr, &sel,
                                     (cudaStream_t)stream);
    if (selector) *selector = sel;
    return rc;
}

}  // extern "C"
