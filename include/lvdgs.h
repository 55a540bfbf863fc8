/*
 * lvdgs.h -- C ABI of the B200-native (sm_100a) differentiable Gaussian-splatting rasterizer.
 *
 * Drop-in boundary for the two native plugins zwk0901/LVD_GS-SLAM calls on every tracking / mapping
 * iteration (README.md:39-44 of the reference: `pip install submodules/simple-knn`,
 * `pip install submodules/diff-gaussian-rasterization`; their sources are in the dropped submodules.zip,
 * /root/reference/.MISSING_LARGE_BLOBS:1).  Each entry point below names the upstream `_C` export it
 * replaces and the in-tree call site that reaches it.
 *
 * Conventions: plain pointers and sizes only; every array pointer is DEVICE memory unless it says "host";
 * float = IEEE binary32; 4x4 matrices are 16 floats in the reference's transposed layout, i.e. the flat
 * array is the column-major math matrix (utils/camera_utils.py:106-120).  `stream` is a cudaStream_t passed
 * as void*.  Functions return 0 on success, non-zero on error (message: lvdgs_last_error()).  No allocation
 * happens inside the library: growable buffers are obtained through the caller's lvdgs_resize_fn, exactly
 * like upstream's `resizeFunctional` over torch uint8 tensors.  The only host synchronisation is the one
 * upstream has too: reading back the instance count R between binning and sorting (hidden behind speculative
 * launches when the caller passes a capacity hint, see lvdgs_rasterize_forward).
 */
#ifndef LVDGS_H
#define LVDGS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LVDGS_TILE 16

/* flags (all off = upstream behaviour) */
#define LVDGS_FLAG_EXACT_PP 1      /* pose Jacobian keeps the principal-point entries Pr[8], Pr[9] */
#define LVDGS_FLAG_OPACITY_GRAD 2  /* propagate dL/d(out_opacity) through the blend backward */
#define LVDGS_FLAG_ACCUMULATE 4    /* backward ADDS the parameter gradients into the caller's buffers (sum over views) */
#define LVDGS_FLAG_POSE_ONLY 8     /* backward produces only dL_dmeans2D (optional) and dL_dtau / dL_dtau_sum; every
                                      parameter-gradient output may be NULL (tracking: only the camera is optimised) */

#define LVDGS_FLAG_GLOBAL_SORT 16  /* sort all (tile | depth) keys with the global onesweep radix sort (as upstream's
                                      cub::DeviceRadixSort) instead of one shared-memory sort per tile segment; the
                                      sorted keys / point list are bit-identical either way.  Must be the same in the
                                      forward and its backward. */

#define LVDGS_FLAG_ZEROED_OUTPUTS 32 /* backward: the caller has zero-filled every gradient output; rows of culled Gaussians are
                                      then left alone and the backward walks only the visible Gaussians */

#define LVDGS_FLAG_ZEROED_SCRATCH 64 /* backward: the caller has zero-filled `scratch` (the first lvdgs_backward_scratch_bytes bytes) since
                                      the last backward that used it; the backward then skips its own memset.  lvdgs.engine clears a
                                      slot's scratch on the forward stream, off the backward stream's critical path */

/* which buffer a resize callback is asked for */
#define LVDGS_BUF_GEOM 0
#define LVDGS_BUF_BINNING 1
#define LVDGS_BUF_IMG 2

typedef struct lvdgs_raster_params {
    int32_t P;            /* Gaussians */
    int32_t sh_degree;    /* active SH degree D (0..3) */
    int32_t sh_coeffs;    /* coefficients stored per Gaussian M >= (D+1)^2; ignored when colors_precomp != NULL */
    int32_t width, height;
    float tan_fovx, tan_fovy;
    float scale_modifier;
    int32_t prefiltered;  /* accepted for signature parity; unused (as upstream) */
    int32_t debug;        /* 1: synchronise and check for CUDA errors after every kernel */
    int32_t flags;        /* LVDGS_FLAG_* */
} lvdgs_raster_params;

/* Must return a device pointer to at least `bytes` bytes, 256-byte aligned, valid until the matching backward
 * has run.  Called once per buffer per forward (twice for LVDGS_BUF_BINNING when a speculative launch has to be
 * repeated: the capacity hint was too small, or a tile list was longer than the previous frame suggested),
 * from the calling thread; the latest pointer per buffer is the live one. */
typedef void *(*lvdgs_resize_fn)(void *user, int32_t which, size_t bytes);

/*
 * A ready-made resize callback for hosts that know the buffer sizes in advance (lvdgs_get_*_layout) and want no call
 * back into their own language per forward: pass lvdgs_static_resize as `resize` and a lvdgs_static_buffers as
 * `resize_user`.  A request that fits capacity[which] is served from base[which]; a larger one goes to `fallback`
 * (NULL: the request fails and the forward returns an error).
 */
typedef struct lvdgs_static_buffers {
    void *base[3];              /* LVDGS_BUF_GEOM / _BINNING / _IMG, 256-byte aligned device pointers */
    size_t capacity[3];         /* bytes available behind each */
    lvdgs_resize_fn fallback;   /* e.g. the host's allocating callback */
    void *fallback_user;
} lvdgs_static_buffers;
void *lvdgs_static_resize(void *user, int32_t which, size_t bytes);

/* Byte offsets of the arrays inside the three opaque buffers (for parity tests and debuggers). */
typedef struct lvdgs_geom_layout {
    size_t depths;         /* float  [P]    view-space z */
    size_t means2D;        /* float4 [P]    pixel centre x,y + half extents hx,hy of the alpha>=1/255 ellipse's bounding box */
    size_t conic_opacity;  /* float4 [P]    conic xx,xy,yy + opacity */
    size_t rgbd;           /* float4 [P]    rgb after SH + clamp, w = depth */
    size_t rect;           /* int16x4 [P]   tile rect min.x,min.y,max.x,max.y */
    size_t tiles_touched;  /* uint32 [P] */
    size_t point_offsets;  /* uint32 [P]    inclusive scan of tiles_touched (written by the key emission) */
    size_t clamped;        /* uint8  [P]    bit c set: channel c clamped at 0 */
    size_t scan_state;     /* uint32 [ceil(P/256)] instances per preprocess block (-> exclusive offsets), then 256 bytes of
                              counters: [0] R, [1] longest tile list, [2] number of visible Gaussians */
    size_t visible_list;   /* uint32 [P]    indices of the Gaussians with tiles_touched > 0, [0, counters[2]) valid */
    size_t total;
} lvdgs_geom_layout;

typedef struct lvdgs_binning_layout {
    size_t keys[2];        /* uint64 [R] x2  (tile << 32 | depth bits); double buffer.  Default (tile-segment) sort:
                              [0] = per-tile segments of (depth bits << 32 | Gaussian) in slot order, [1] = sorted keys */
    size_t vals[2];        /* uint32 [R] x2  Gaussian index */
    size_t sort_ws;        /* onesweep histograms + look-back state */
    size_t sorted_sel;     /* int32: which of the two key/val buffers holds the sorted result */
    size_t total;
} lvdgs_binning_layout;

typedef struct lvdgs_img_layout {
    size_t final_T;        /* float  [H*W] */
    size_t n_contrib;      /* uint32 [H*W] */
    size_t ranges;         /* uint2  [tiles] */
    size_t tile_order;     /* uint32 [tiles] tile ids, heaviest lists first: launch order of the blend kernels */
    size_t tile_grid;      /* int32  [(gy+1)*(gx+1)] difference array -> per-tile instance counts */
    size_t sort_hist;      /* uint32 [8][256] exclusive-scanned digit histograms of the sort keys (LVDGS_FLAG_GLOBAL_SORT) */
    size_t tile_cursor;    /* uint32 [tiles][8] (one per 32-byte sector) instances emitted so far into each tile's segment */
    size_t total;
} lvdgs_img_layout;

int lvdgs_version(void);
const char *lvdgs_last_error(void);
int lvdgs_set_device(int device);
/* kernels launched by this library since the last reset (bench.py's gpu_launches) */
int64_t lvdgs_launch_count(void);
void lvdgs_reset_launch_count(void);
/* forwards of the calling thread whose speculatively launched tail had to be repeated (capacity hint too small, or a tile
 * list longer than the recent frames of the same device and image size suggested) -- a cost counter, never a correctness
 * matter */
int64_t lvdgs_tail_rerun_count(void);

/*
 * Per-launch device timing for bench.py's roofline: between begin and end every kernel launch of this library is
 * followed by a CUDA event on the launching stream.  lvdgs_profile_end synchronises the stream and returns the
 * number of entries n (or -1): ms[i] = time between the events before and after launch i, names = n
 * newline-separated kernel names.  Not thread-safe; never enable inside a timed region.
 */
int lvdgs_profile_begin(void *stream);
int lvdgs_profile_end(void *stream, char *names, size_t names_bytes, float *ms, int32_t max_entries);

int lvdgs_get_geom_layout(int32_t P, lvdgs_geom_layout *out);
int lvdgs_get_binning_layout(int64_t R, lvdgs_binning_layout *out);
int lvdgs_get_img_layout(int32_t width, int32_t height, lvdgs_img_layout *out);

/*
 * Replaces `_C.rasterize_gaussians` (upstream rasterize_points.cu: RasterizeGaussiansCUDA), reached from
 * gaussian_renderer.render at utils/slam_frontend.py:1493, utils/slam_backend.py:98,184,277,407,
 * utils/eval_utils_0806.py:215 and render_with_custom_resolution at utils/init_pose.py:145.
 *   means3D [P,3]; opacities [P]; exactly one of {shs [P,M,3], colors_precomp [P,3]};
 *   exactly one of {scales [P,3] + rotations [P,4], cov3D_precomp [P,6]}; background [3]; campos [3].
 * Outputs: out_color [3,H,W], radii [P] int32, out_depth [H,W], out_opacity [H,W], n_touched [P] int32,
 *   *num_rendered (host) = R, the number of (tile, Gaussian) instances; *binning_capacity (host) = the instance
 *   capacity the binning buffer was laid out for (pass both to the backward).
 * capacity_hint: 0 = size the binning buffer exactly, after reading R back (upstream's behaviour: the device idles
 *   while the host reads R and launches the rest).  > 0 = speculative mode: the binning buffer is requested for
 *   `capacity_hint` instances and the whole remainder of the forward is queued BEFORE the host waits for R; if
 *   R > capacity_hint the remainder is re-run with an exactly sized buffer (resize is then called a second time
 *   for LVDGS_BUF_BINNING).  Results are identical in both modes.
 */
int lvdgs_rasterize_forward(const lvdgs_raster_params *prm, const float *background, const float *means3D,
                            const float *colors_precomp, const float *opacities, const float *scales,
                            const float *rotations, const float *cov3D_precomp, const float *viewmatrix,
                            const float *projmatrix, const float *projmatrix_raw, const float *shs,
                            const float *campos, lvdgs_resize_fn resize, void *resize_user, int64_t capacity_hint,
                            float *out_color, int32_t *radii, float *out_depth, float *out_opacity,
                            int32_t *n_touched, int64_t *num_rendered, int64_t *binning_capacity, void *stream);

/* cudaMemsetAsync(ptr, 0, bytes, stream) for callers without a CUDA runtime binding (used with LVDGS_FLAG_ZEROED_SCRATCH). */
int lvdgs_zero_async(void *ptr, size_t bytes, void *stream);

/* Device scratch needed by lvdgs_rasterize_backward for P Gaussians and R instances. */
size_t lvdgs_backward_scratch_bytes(int32_t P, int64_t R);

/*
 * Replaces `_C.rasterize_gaussians_backward` (upstream RasterizeGaussiansBackwardCUDA), reached through
 * loss.backward() at utils/slam_frontend.py:1517 and utils/slam_backend.py:120,306,457.
 *   dL_dout_color [3,H,W]; dL_dout_depth [H,W] or NULL; dL_dout_opacity [H,W] or NULL (used only with
 *   LVDGS_FLAG_OPACITY_GRAD; upstream drops it).  geom/binning/img buffers and R are those of the forward.
 * Outputs (every element written, no pre-zeroing needed -- except with LVDGS_FLAG_ACCUMULATE, where the parameter
 *   gradients dL_dcolors/opacity/means3D/cov3D/sh/scales/rots are added to the buffers' contents): dL_dmeans2D [P,3]
 *   (z = 0; NULL ok), dL_dcolors [P,3] (NULL ok unless colors_precomp), dL_dopacity [P], dL_dmeans3D [P,3],
 *   dL_dcov3D [P,6] (NULL ok unless cov3D_precomp), dL_dsh [P,M,3] (NULL ok with colors_precomp),
 *   dL_dscales [P,3], dL_drots [P,4] (NULL ok with cov3D_precomp), dL_dtau [P,6] = (rho, theta) per Gaussian
 *   (NULL ok), dL_dtau_sum [6] = the sum over Gaussians that upstream forms in Python
 *   (`grad_tau.view(-1,6).sum(0)`) (NULL ok).
 */
int lvdgs_rasterize_backward(const lvdgs_raster_params *prm, const float *background, const float *means3D,
                             const int32_t *radii, const float *colors_precomp, const float *opacities,
                             const float *scales, const float *rotations, const float *cov3D_precomp,
                             const float *viewmatrix, const float *projmatrix, const float *projmatrix_raw,
                             const float *dL_dout_color, const float *dL_dout_depth, const float *dL_dout_opacity,
                             const float *shs, const float *campos, const void *geom_buffer, int64_t R,
                             int64_t binning_capacity, const void *binning_buffer, const void *img_buffer, void *scratch,
                             size_t scratch_bytes, float *dL_dmeans2D, float *dL_dcolors, float *dL_dopacity,
                             float *dL_dmeans3D, float *dL_dcov3D, float *dL_dsh, float *dL_dscales,
                             float *dL_drots, float *dL_dtau, float *dL_dtau_sum, void *stream);

/* Replaces `_C.mark_visible` (GaussianRasterizer.markVisible): present[i] = view-space z > 0.2. */
int lvdgs_mark_visible(int32_t P, const float *means3D, const float *viewmatrix, const float *projmatrix,
                       uint8_t *present, void *stream);

/*
 * Replaces `simple_knn._C.distCUDA2` (reached from GaussianModel.extend_from_pcd_seq, utils/slam_backend.py:75-78):
 * mean_dists[i] = mean squared distance from points[i] to its 3 nearest neighbours (exact).
 */
size_t lvdgs_dist2_workspace_bytes(int32_t P);
int lvdgs_dist2(int32_t P, const float *points, float *mean_dists, void *workspace, size_t workspace_bytes,
                void *stream);

/*
 * Fused Adam update of a flat float32 parameter block with per-group learning rates (group k covers elements
 * [group_end[k-1], group_end[k]); group_end and lr are HOST arrays of `groups` <= 8 entries; step >= 1 is the
 * 1-based iteration for the bias correction).  Replaces the torch.optim.Adam.step the reference runs after every
 * mapping iteration (utils/slam_backend.py:378-380); used after the gradient all-reduce of the keyframe-sharded
 * mapping step so that every rank applies the identical update.
 */
int lvdgs_adam_step(int64_t n, float *params, const float *grads, float *exp_avg, float *exp_avg_sq, int32_t groups,
                    const int64_t *group_end, const float *lr, double beta1, double beta2, double eps, int32_t step,
                    void *stream);

/*
 * The exchange step of the keyframe-sharded mapping iteration as ONE kernel over NVLink peer memory (the
 * single-GPU reference calls GaussianModel.optimizer.step(), utils/slam_backend.py:378-380).  grad_ptrs / param_ptrs /
 * act_ptrs: HOST arrays of `world` device pointers, entry r = rank r's gradient block / raw parameter block / activated
 * block (opacity | scales | rotations) mapped into this process (peer access); act_ptrs NULL = the block holds no raw
 * parameters.  For elements [lo, hi) of the block (this rank's slice, float4 aligned) the kernel sums the gradient over
 * the ranks, applies the activation chain rule, Adam (moments exp_avg / exp_avg_sq are LOCAL full-size arrays, only the
 * slice is touched) and stores the new raw values and activations into every rank's blocks.  group_end / lr as in
 * lvdgs_adam_step (5 groups: means3D, shs, opacity, scales, rotations; every group starts 16-byte aligned);
 * act_offsets[3]: starts of opacity / scales / rotations inside the activated block (floats), act_total its size.
 * mc_grad / mc_param / mc_act: NVSwitch MULTICAST mappings of the three blocks (torch symmetric memory's multicast_ptr), or
 * NULL.  With all of them the sum over the ranks is one in-switch `multimem.ld_reduce` and every result is sent once and
 * replicated by the switch (`multimem.st`); without, the kernel loads from / stores to every peer itself.
 * act_mode: 0 = the block holds no raw parameters (no chain rule, no activations; act_ptrs may be NULL); 1 = raw block, the
 * kernel stores the activations of its slice into every rank's activated block; 2 = raw block, activations are NOT
 * stored -- the caller runs lvdgs_gaussian_activate over the whole block after the closing barrier (8 of the 14 floats per
 * Gaussian less over NVLink for one local pass over HBM).
 * The caller provides the two cross-rank barriers around the launch.
 */
int lvdgs_exchange_adam(int32_t world, int32_t rank, const float *const *grad_ptrs, float *const *param_ptrs, float *const *act_ptrs,
                        int64_t lo, int64_t hi, float *exp_avg, float *exp_avg_sq, int32_t groups, const int64_t *group_end,
                        const float *lr, const int64_t *act_offsets, int64_t act_total, double beta1, double beta2, double eps,
                        int32_t step, const float *mc_grad, float *mc_param, float *mc_act, int32_t act_mode, void *stream);

/*
 * The binning sort on its own (stable LSD radix sort of u64 keys with u32 values over key bits
 * [0, end_bit)), exposed for parity tests and for the comparison against cub::DeviceRadixSort, the
 * library call upstream makes.  keys/vals: two buffers each of n elements; input in [0];
 * *selector (host) receives the buffer index that holds the sorted result.
 */
size_t lvdgs_sort_workspace_bytes(int64_t n);
int lvdgs_sort_pairs(int64_t n, uint64_t *keys0, uint64_t *keys1, uint32_t *vals0, uint32_t *vals1,
                     int32_t end_bit, void *workspace, size_t workspace_bytes, int32_t *selector, void *stream);
size_t lvdgs_cub_sort_workspace_bytes(int64_t n, int32_t end_bit);
int lvdgs_cub_sort_pairs(int64_t n, const uint64_t *keys_in, uint64_t *keys_out, const uint32_t *vals_in,
                         uint32_t *vals_out, int32_t end_bit, void *workspace, size_t workspace_bytes,
                         void *stream);

/*
 * Measurement aid for bench.py (not a product path): one launch of `blocks` x 256 threads, each running iters x 16
 * independent register-only operations.  mode 0: scalar FFMA, mode 1: packed FFMA2 (two fp32 FMAs per instruction);
 * modes 2..7: instruction-mix probes (FMUL2, FADD2, FFMA2 + FFMA, FFMA2 + select, select alone, FFMA + select; see
 * csrc/peak.cu and scripts/pipe_probe.py).  *fmas (host) receives the number of FMAs (modes 0, 1, 4) or instructions of the
 * launch; time it with CUDA events -> the device's measured FP32 FMA-pipe peak, the denominator of the blend kernels'
 * roofline fraction.  out: any device float (never written).
 */
int lvdgs_fp32_peak(int32_t blocks, int32_t iters, int32_t mode, float *out, double *fmas, void *stream);

/* ------------------------------------------------------------------------------------------------------------
 * The callers either side of the rasterizer (SURVEY.md section 8f "next" rows).  Same conventions as above.
 * ------------------------------------------------------------------------------------------------------------ */

#define LVDGS_LOSS_OPACITY_WEIGHT 1       /* rgb term weighted by the rendered opacity (tracking: utils/slam_utils.py:60) */
#define LVDGS_LOSS_DEPTH_NEEDS_OPAQUE 2   /* depth term only where opacity > 0.95 (tracking rgbd: utils/slam_utils.py:74) */

/*
 * Row N3.  Replaces the torch expression graphs of get_loss_tracking / get_loss_tracking_rgb / get_loss_tracking_rgbd
 * (utils/slam_utils.py:42-83) and get_loss_mapping / _rgb / _rgbd (:86-121) together with their autograd backward:
 *   loss = w_rgb * mean_{c,p} o | m (exp(a) I_c + b) - m gt_c |  +  w_depth * mean_p | md D - md gtD |
 *   m = (sum_c gt_c > rgb_boundary_threshold) * grad_mask,  o = opacity or 1,  md = (gtD > 0.01) [* (opacity > 0.95)].
 * color [3,H,W], depth / opacity / gt_depth / grad_mask [H,W] (gt_depth NULL or w_depth == 0: no depth term; grad_mask,
 * opacity NULL: ones), exposure = device pointer to {a, b} or NULL (a = b = 0).
 * Outputs: g_color [3,H,W] = dL/dI, g_depth [H,W] (NULL ok), g_opacity [H,W] (NULL ok),
 *   out[4] (device) = {loss, dL/da, dL/db, 0}.  Deterministic (fixed summation order).
 */
size_t lvdgs_fused_loss_workspace_bytes(void);
int lvdgs_fused_loss(int32_t width, int32_t height, const float *color, const float *depth, const float *opacity,
                     const float *gt_color, const float *gt_depth, const float *grad_mask, const float *exposure,
                     float rgb_boundary_threshold, float w_rgb, float w_depth, int32_t flags, float *g_color,
                     float *g_depth, float *g_opacity, float *out, void *workspace, size_t workspace_bytes, void *stream);

/*
 * Row N3, the masked mapping loss LVD-GS runs when a keyframe carries a static mask (utils/slam_backend.py:199-261):
 * dynamic pixels (static_mask == 0) are painted with `background` in both the render and the ground truth, then
 *   loss = (1 - lambda_dssim) * l1_loss(mi, mg) + lambda_dssim * (1 - ssim(mi, mg))
 *        + depth_lambda * mean_{static & mono > 0 & D > 0} |D - mono|              (dropped when that set is empty)
 * with gaussian_splatting.utils.loss_utils' l1_loss / ssim (11x11 Gaussian window, sigma 1.5, zero padding).
 * image / gt_image [3,H,W]; static_mask uint8 [H,W] or NULL (all static); background [3]; depth / mono_depth [H,W] or NULL.
 * Outputs: g_image [3,H,W] = dL/dimage, g_depth [H,W] (NULL ok), out[8] (device) = {loss, mean SSIM, mean |mi - mg|,
 * depth mean, depth pixel count, ...}.  Deterministic.  workspace: zero-filled before its first use.
 */
size_t lvdgs_masked_ssim_loss_workspace_bytes(int32_t width, int32_t height);
int lvdgs_masked_ssim_loss(int32_t width, int32_t height, const float *image, const float *gt_image, const uint8_t *static_mask,
                           const float *background, const float *depth, const float *mono_depth, float lambda_dssim,
                           float depth_lambda, float *g_image, float *g_depth, float *out, void *workspace, size_t workspace_bytes,
                           void *stream);

/*
 * Row N4.  Keyframe covisibility from per-Gaussian visibility (n_touched > 0), replacing the logical_and / logical_or
 * + count_nonzero chains of utils/slam_frontend.py:1598-1603,1631-1639: out[4] (device, uint64) =
 * {|a|, |b|, |a and b|, |a or b|}; an element is visible when non-zero; elem_bytes in {1, 4, 8} (bool / int32 / int64).
 */
int lvdgs_covis_counts(int64_t n, const void *a, const void *b, int32_t elem_bytes, uint64_t *out, void *stream);
/* n_obs[i] = number of the K visibility arrays (device array of K device pointers) that are non-zero at i; replaces the
 * CPU accumulation `n_obs += visibility.cpu()` of utils/slam_backend.py:322-325. */
int lvdgs_n_obs(int64_t n, int32_t K, const void *const *masks, int32_t elem_bytes, int32_t *n_obs, void *stream);

/*
 * Row N1.  Stable compaction of up to 16 row-major float arrays by one keep mask (uint8, non-zero = keep): the
 * parameter / Adam-moment surgery of GaussianModel.prune_points (callers utils/slam_backend.py:128-145,322-339).
 * lvdgs_compact_count scans the mask and leaves the number of kept rows in *count_dev (a device uint32 inside the
 * workspace) -- read it back to size the destinations; lvdgs_compact_move (same mask, same workspace, any number of
 * calls) writes the kept rows of src[k] ([n, widths[k]]) to dst[k] in their original order.  dst must not alias src.
 */
size_t lvdgs_compact_workspace_bytes(int64_t n);
int lvdgs_compact_count(int64_t n, const uint8_t *keep, void *workspace, size_t workspace_bytes, uint32_t **count_dev,
                        void *stream);
int lvdgs_compact_move(int64_t n, const uint8_t *keep, const void *workspace, int32_t n_arrays,
                       const float *const *src, float *const *dst, const int32_t *widths, void *stream);

/*
 * Row N1, densification half: dst[k][j, :] = src[k][idx[j], :] for up to 16 row-major float arrays ([n_src_rows,
 * widths[k]]) in one launch -- the row copies behind GaussianModel.densify_and_clone / densify_and_split
 * (reached from utils/slam_backend.py:359-376), whose new rows are appended to every parameter tensor.  idx: device
 * int64 [n_idx]; out-of-range indices leave the destination row untouched.  dst may not overlap the rows it reads.
 */
int lvdgs_gather_rows(int64_t n_idx, const int64_t *idx, int64_t n_src_rows, int32_t n_arrays, const float *const *src,
                      float *const *dst, const int32_t *widths, void *stream);

/*
 * Row N1, parametrisation.  GaussianModel optimises RAW parameters and hands the rasterizer their activations
 * (get_opacity = sigmoid(_opacity), get_scaling = exp(_scaling), get_rotation = normalize(_rotation); used by render() at
 * utils/slam_backend.py:98,184,277,407).  lvdgs_gaussian_activate: raw [P], [P,3], [P,4] -> activated arrays of the same
 * shapes.  lvdgs_gaussian_activation_backward: turns the gradients with respect to the ACTIVATED values (the outputs of
 * lvdgs_rasterize_backward, possibly summed over views) into gradients with respect to the RAW parameters, in place --
 * the chain rule torch autograd applies in the reference.  rotations / raw_rotations / g_rotations 16-byte aligned.
 */
int lvdgs_gaussian_activate(int64_t P, const float *raw_opacity, const float *raw_scales, const float *raw_rotations, float *opacity,
                            float *scales, float *rotations, void *stream);
int lvdgs_gaussian_activation_backward(int64_t P, const float *opacity, const float *scales, const float *rotations,
                                       const float *raw_rotations, float *g_opacity, float *g_scales, float *g_rotations, void *stream);

/*
 * Rows a15 / a16: the tail of one tracking iteration on the device.  lvdgs_pose_state is the camera's device-resident
 * block; view / proj / campos are in the layout lvdgs_rasterize_* read (pass pointers into the block), so a tracking
 * loop needs no host arithmetic between iterations.  lvdgs_pose_step = torch.optim.Adam.step on (cam_rot_delta,
 * cam_trans_delta, exposure_a, exposure_b) with learning rates (lr_rot, lr_trans, lr_exposure) followed by update_pose
 * (utils/slam_frontend.py:1466-1521, utils/pose_utils.py:70-87): T_w2c <- SE3_exp([trans_delta; rot_delta]) T_w2c, the
 * three camera matrices refreshed, converged = |tau| < converged_threshold.  g_tau = (rho[3], theta[3]) as produced by
 * lvdgs_rasterize_backward's dL_dtau_sum; g_exposure = {dL/da, dL/db} (lvdgs_fused_loss out + 1) or NULL (exposure
 * fixed).  `step` is the 1-based iteration of this frame's optimiser (Adam bias correction); the host initialises the block
 * (R, T, proj_raw, exposure; moments zero) and the matrices are valid after the first lvdgs_pose_step or when the host
 * fills them too.
 */
typedef struct lvdgs_pose_state {
    float view[16];       /* world_view_transform = [R T; 0 1]^T, flat row-major (utils/camera_utils.py:106-108) */
    float proj[16];       /* full_proj_transform = world_view_transform @ projection_matrix (:110-116) */
    float proj_raw[16];   /* projection_matrix (P^T), constant per camera */
    float campos[4];      /* camera_center (:118-120), w unused */
    float R[9];           /* world -> camera rotation, row-major */
    float T[3];
    float exposure[4];    /* exposure_a, exposure_b, unused, unused */
    float adam_m[8];      /* first / second moments of (rot_delta[3], trans_delta[3], exposure_a, exposure_b) */
    float adam_v[8];
    int32_t step;         /* pose steps taken (device-side count) */
    int32_t converged;    /* |tau| < threshold at the last step */
    float tau_norm;
    float pad;
} lvdgs_pose_state;

int lvdgs_pose_step(lvdgs_pose_state *state, const float *g_tau, const float *g_exposure, float lr_rot, float lr_trans,
                    float lr_exposure, double beta1, double beta2, double eps, int32_t step, float converged_threshold,
                    void *stream);

#ifdef __cplusplus
}
#endif
#endif /* LVDGS_H */
