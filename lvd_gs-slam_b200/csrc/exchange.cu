// exchange.cu -- the exchange step of the keyframe-sharded mapping iteration as ONE kernel over NVLink peer memory
// (the reference has one GPU and calls GaussianModel.optimizer.step(), utils/slam_backend.py:378-380).
//
// Every rank holds the replicated RAW parameter block, its activated copies, and the gradient block its views
// accumulated.  Rank r owns elements [lo, hi) of the block (1/world, ZeRO-1).  For its slice the kernel
//   1. loads the slice of EVERY rank's gradient block straight from that rank's memory (peer loads over NVLink; eight
//      128-bit loads in flight per thread) and sums them in rank order                          -- the reduce-scatter,
//   2. takes the sum back through the activations to the raw parameters (sigmoid / exp / normalize chain rule; the
//      activations are recomputed from the raw value with the expressions of gaussian_activate_kernel, so the result
//      is bit-identical to lvdgs_gaussian_activation_backward on the summed block -- the chain rule is linear in g),
//   3. applies Adam with the group's learning rate to its slice of the moments and raw parameters -- the optimiser step,
//   4. stores the new raw values into every rank's parameter block (peer stores) -- the all-gather -- and, with act_mode 1,
//      their activations into every rank's activated block too; act_mode 2 (lvdgs.mapping's default) leaves the activations
//      to one local lvdgs_gaussian_activate pass after the closing barrier: 8 of the 14 floats per Gaussian less on the wire.
// Bytes over NVLink per rank and step at world w: (w-1)/w of the gradient block in, (w-1)/w x (parameters + activated
// groups) out, the same as reduce-scatter + all-gather, but there is one launch instead of five (chain rule, NCCL
// reduce-scatter, Adam, NCCL all-gather, activate) and no staging copy.  The caller brackets the launch with two
// cross-rank barriers (all gradients complete before; all stores landed and all gradient slices consumed after).
//
// With NVSwitch multicast mappings of the three blocks (mc_* != nullptr; torch symmetric memory's multicast_ptr) steps 1
// and 4 run INSIDE the switch: one `multimem.ld_reduce.add.v4.f32` returns the sum of the slice over all ranks (the
// switch pulls and adds the eight copies; 1/w of the unicast read traffic arrives at this GPU), and one `multimem.st`
// per value is replicated by the switch to every rank (this GPU sends each byte once instead of w-1 times): per rank
// and step block/w bytes in and (parameters + activated groups)/w bytes out, against (w-1) times that for peer loads
// and stores.  The order of the in-switch sum is the switch's; replicas stay bit-identical because only the slice's
// owner reduces and every rank receives the owner's result.
//
// A world of one is the single-GPU update: chain rule + Adam + activations in one launch instead of three.  Measured
// (DESIGN.md section 8): 8 B200s, 500 k Gaussians: 0.103-0.111 ms for the whole exchange step (barrier 12 us, kernel 64-75 us,
// barrier 17-20 us, local activations + gradient clear 17 us) against 0.173 ms for the NCCL sequence; the kernel time
// is set by every GPU serving (w-1)/w of its gradient block to the switch, not by the arithmetic.
#include "common.cuh"
#include <cmath>

namespace lvdgs {

constexpr int EX_MAX_WORLD = 16;

struct ExchangeArgs {
    const float *grad[EX_MAX_WORLD];     // every rank's gradient block (peer-mapped device pointers)
    float *param[EX_MAX_WORLD];          // every rank's raw parameter block
    float *act[EX_MAX_WORLD];            // every rank's activated block (opacity | scales | rotations), or nullptr (no activations)
    const float *mc_grad;                // NVSwitch multicast mappings of the same three blocks, or nullptr (peer loads / stores)
    float *mc_param, *mc_act;
    int world;
    int raw, store_act;                  // block holds raw parameters (chain rule applies); activations are stored by this kernel
    int64_t lo4, hi4;                    // this rank's slice in float4 units
    int64_t group_end4[8];               // exclusive end of each parameter group in float4 units (means3D, shs, opacity, scales, rotations)
    float lr[8];
    int groups;
    int64_t off_opacity4, off_scales4, off_rot4;            // group starts inside the parameter block, float4 units
    int64_t act_opacity4, act_scales4, act_rot4;            // the same groups inside the activated block
    int64_t act_total4;                                     // size of the activated block (the block's tail padding has no activation)
    float b1, b2, omb1, omb2, eps, inv_bc1, inv_sqrt_bc2;
};

__device__ __forceinline__ float4 multimem_sum4(const float4 *p) {        // in-switch reduction over every rank's copy
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void multimem_store4(float4 *p, float4 v) {     // one store, replicated by the switch to every rank
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ float adam_update(float p, float g, float &m, float &v, float lr, const ExchangeArgs &a) {
    const float mi = a.b1 * m + a.omb1 * g;                  // same expressions as adam_step_kernel (adam.cu)
    const float vi = a.b2 * v + a.omb2 * g * g;
    m = mi; v = vi;
    const float denom = sqrtf(vi) * a.inv_sqrt_bc2 + a.eps;
    return p - lr * a.inv_bc1 * mi / denom;
}

// MAXW: compile-time bound of the world size (the peer-load path keeps one 128-bit load per rank in flight: the register
// footprint follows MAXW -- 116 registers and 2 CTAs / SM for the generic 16, ~48 for the in-switch path); MC: the three
// blocks have multicast mappings
template <int MAXW, bool MC>
__global__ void __launch_bounds__(256, MC ? 4 : (MAXW <= 4 ? 4 : 2)) exchange_adam_kernel(float4 *__restrict__ exp_avg, float4 *__restrict__ exp_avg_sq, const ExchangeArgs a) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = a.lo4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.hi4; i += stride) {
        // 1. reduce: this element's gradient from every rank -- in the switch, or summed here in rank order
        float4 g;
        if (MC) {
            g = multimem_sum4(reinterpret_cast<const float4 *>(a.mc_grad) + i);
        } else {
            float4 gr[MAXW];
#pragma unroll
            for (int r = 0; r < MAXW; ++r)
                if (r < a.world) gr[r] = __ldcg(reinterpret_cast<const float4 *>(a.grad[r]) + i);      // L2 only: the line is remote and read once
            g = gr[0];
#pragma unroll
            for (int r = 1; r < MAXW; ++r)
                if (r < a.world) { g.x += gr[r].x; g.y += gr[r].y; g.z += gr[r].z; g.w += gr[r].w; }
        }
        int grp = 0;
#pragma unroll
        for (int k = 1; k < 8; ++k)
            if (k < a.groups && i >= a.group_end4[k - 1]) grp = k;
        const float lr = a.lr[grp];
        // every rank's parameter block is identical: read the local one (slot `world` = this rank's own block)
        float4 p = reinterpret_cast<const float4 *>(a.param[a.world])[i];
        // 2. chain rule through the activation of this group (act != nullptr: the block holds raw parameters)
        const bool raw = a.raw != 0;
        if (raw && grp == 2) {              // opacity = sigmoid(raw)
            const float o0 = 1.f / (1.f + expf(-p.x)), o1 = 1.f / (1.f + expf(-p.y)), o2 = 1.f / (1.f + expf(-p.z)), o3 = 1.f / (1.f + expf(-p.w));
            g.x *= o0 * (1.f - o0); g.y *= o1 * (1.f - o1); g.z *= o2 * (1.f - o2); g.w *= o3 * (1.f - o3);
        } else if (raw && grp == 3) {       // scale = exp(raw)
            g.x *= expf(p.x); g.y *= expf(p.y); g.z *= expf(p.z); g.w *= expf(p.w);
        } else if (raw && grp == 4) {       // rotation = raw / |raw| (one quaternion per float4)
            const float n = fmaxf(sqrtf(p.x * p.x + p.y * p.y + p.z * p.z + p.w * p.w), 1e-12f);
            const float4 q = make_float4(p.x / n, p.y / n, p.z / n, p.w / n);
            const float inv = 1.f / n;
            const float d = q.x * g.x + q.y * g.y + q.z * g.z + q.w * g.w;
            g = make_float4((g.x - q.x * d) * inv, (g.y - q.y * d) * inv, (g.z - q.z * d) * inv, (g.w - q.w * d) * inv);
        }
        // 3. Adam on this rank's slice of the moments
        float4 m = exp_avg[i], v = exp_avg_sq[i];
        p.x = adam_update(p.x, g.x, m.x, v.x, lr, a);
        p.y = adam_update(p.y, g.y, m.y, v.y, lr, a);
        p.z = adam_update(p.z, g.z, m.z, v.z, lr, a);
        p.w = adam_update(p.w, g.w, m.w, v.w, lr, a);
        exp_avg[i] = m; exp_avg_sq[i] = v;
        // 4. all-gather + activate: the new raw values and their activations go to every rank
        float4 av = p;
        int64_t ai = -1;
        if (!a.store_act) {
            // the caller activates the whole block locally after the closing barrier: 8 of the 14 floats per Gaussian less on the wire
        } else if (raw && grp == 2) {
            av = make_float4(1.f / (1.f + expf(-p.x)), 1.f / (1.f + expf(-p.y)), 1.f / (1.f + expf(-p.z)), 1.f / (1.f + expf(-p.w)));
            ai = a.act_opacity4 + (i - a.off_opacity4);
        } else if (raw && grp == 3) {
            av = make_float4(expf(p.x), expf(p.y), expf(p.z), expf(p.w));
            ai = a.act_scales4 + (i - a.off_scales4);
        } else if (raw && grp == 4) {
            const float n = fmaxf(sqrtf(p.x * p.x + p.y * p.y + p.z * p.z + p.w * p.w), 1e-12f);
            av = make_float4(p.x / n, p.y / n, p.z / n, p.w / n);
            ai = a.act_rot4 + (i - a.off_rot4);
        }
        if (MC) {
            multimem_store4(reinterpret_cast<float4 *>(a.mc_param) + i, p);
            if (ai >= 0 && ai < a.act_total4) multimem_store4(reinterpret_cast<float4 *>(a.mc_act) + ai, av);
        } else {
#pragma unroll
            for (int r = 0; r < MAXW; ++r)
                if (r < a.world) {
                    __stcg(reinterpret_cast<float4 *>(a.param[r]) + i, p);
                    if (ai >= 0 && ai < a.act_total4) __stcg(reinterpret_cast<float4 *>(a.act[r]) + ai, av);
                }
        }
    }
}

int launch_exchange_adam(int world, int rank, const float *const *grad_ptrs, float *const *param_ptrs, float *const *act_ptrs,
                         int64_t lo, int64_t hi, float *exp_avg, float *exp_avg_sq, int groups, const int64_t *group_end,
                         const float *lr, const int64_t *act_offsets, int64_t act_total, double beta1, double beta2, double eps, int step,
                         const float *mc_grad, float *mc_param, float *mc_act, int act_mode, cudaStream_t s) {
    if (world < 1 || world >= EX_MAX_WORLD) { set_error("exchange: world must be 1..%d", EX_MAX_WORLD - 1); return 1; }
    if (rank < 0 || rank >= world) { set_error("exchange: bad rank"); return 1; }
    if (groups != 5) { set_error("exchange: the block has 5 parameter groups (means3D, shs, opacity, scales, rotations)"); return 1; }
    if ((lo & 3) || (hi & 3) || hi < lo) { set_error("exchange: the slice must be float4-aligned"); return 1; }
    if (hi == lo) return 0;
    ExchangeArgs a{};
    a.world = world;
    if (act_mode < 0 || act_mode > 2 || (act_mode == 1 && !act_ptrs)) { set_error("exchange: bad act_mode"); return 1; }
    a.raw = act_mode != 0; a.store_act = act_mode == 1;
    for (int r = 0; r < world; ++r) {
        a.grad[r] = grad_ptrs[r]; a.param[r] = param_ptrs[r]; a.act[r] = act_ptrs ? act_ptrs[r] : nullptr;
        if (!a.grad[r] || !a.param[r]) { set_error("exchange: NULL peer pointer"); return 1; }
    }
    a.param[world] = param_ptrs[rank];            // this rank's own block, for the local read
    // all three or none: a block without a multicast mapping keeps the whole step on peer loads / stores
    const bool mc = mc_grad && mc_param && (mc_act || act_mode != 1);
    a.mc_grad = mc ? mc_grad : nullptr; a.mc_param = mc ? mc_param : nullptr; a.mc_act = mc ? mc_act : nullptr;
    a.lo4 = lo / 4; a.hi4 = hi / 4;
    a.groups = groups;
    int64_t start = 0;
    int64_t starts[8] = {0};
    for (int k = 0; k < 8; ++k) {
        const int64_t e = k < groups ? group_end[k] : group_end[groups - 1];
        if (e & 3) { set_error("exchange: parameter groups must start on 16-byte boundaries"); return 1; }
        starts[k] = start;
        a.group_end4[k] = e / 4; a.lr[k] = k < groups ? lr[k] : 0.f;
        start = e;
    }
    a.off_opacity4 = starts[2] / 4; a.off_scales4 = starts[3] / 4; a.off_rot4 = starts[4] / 4;
    if (act_mode == 1) {
        if (!act_offsets) { set_error("exchange: act_offsets missing"); return 1; }
        for (int k = 0; k < 3; ++k)
            if (act_offsets[k] & 3) { set_error("exchange: activated groups must start on 16-byte boundaries"); return 1; }
        a.act_opacity4 = act_offsets[0] / 4; a.act_scales4 = act_offsets[1] / 4; a.act_rot4 = act_offsets[2] / 4;
        a.act_total4 = act_total / 4;
    }
    a.b1 = (float)beta1; a.b2 = (float)beta2; a.eps = (float)eps;
    a.omb1 = (float)(1.0 - beta1); a.omb2 = (float)(1.0 - beta2);
    a.inv_bc1 = (float)(1.0 / (1.0 - pow(beta1, (double)step)));
    a.inv_sqrt_bc2 = (float)(1.0 / sqrt(1.0 - pow(beta2, (double)step)));
    const int64_t n4 = a.hi4 - a.lo4;
    const int blocks = (int)min((int64_t)148 * 8, (n4 + 255) / 256);
    LVDGS_PRE(s);
    float4 *m4 = reinterpret_cast<float4 *>(exp_avg), *v4 = reinterpret_cast<float4 *>(exp_avg_sq);
    if (mc) exchange_adam_kernel<1, true><<<blocks, 256, 0, s>>>(m4, v4, a);
    else if (world == 1) exchange_adam_kernel<1, false><<<blocks, 256, 0, s>>>(m4, v4, a);
    else if (world <= 2) exchange_adam_kernel<2, false><<<blocks, 256, 0, s>>>(m4, v4, a);
    else if (world <= 4) exchange_adam_kernel<4, false><<<blocks, 256, 0, s>>>(m4, v4, a);
    else if (world <= 8) exchange_adam_kernel<8, false><<<blocks, 256, 0, s>>>(m4, v4, a);
    else exchange_adam_kernel<EX_MAX_WORLD, false><<<blocks, 256, 0, s>>>(m4, v4, a);
    LVDGS_LAUNCHED(s, "exchange_adam");
    return 0;
}

}  // namespace lvdgs
