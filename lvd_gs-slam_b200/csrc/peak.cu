// peak.cu -- FP32 FMA-pipe peak of this device, measured: the denominator of the blend kernels' roofline fraction
// (MEASURED_PEAKS.json covers HBM and bf16 tensor throughput only).  Each thread runs `iters` rounds of independent
// operations on register operands; bench.py times the launch with CUDA events and divides the work by the duration.
// mode 0: 16 scalar FFMA per round; mode 1: 16 packed FFMA2 (two fp32 FMAs per instruction); the other modes are
// instruction-mix probes used by scripts/pipe_probe.py (how packed and scalar FP32 share the FMA pipes, what an ALU-pipe
// select costs beside them) -- the numbers DESIGN.md's blend-kernel ceiling analysis rests on.
// Not on any product call path.
#include "common.cuh"

namespace lvdgs {

__device__ __forceinline__ float fsel_asm(float a, float b, float c) {      // ALU pipe, two instructions: a > c ? a : b
    float d;
    asm volatile("{ .reg .pred p; setp.gt.f32 p, %1, %3; selp.f32 %0, %1, %2, p; }" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

template <int MODE>
__global__ void __launch_bounds__(256) fp32_peak_kernel(int iters, float seed, float *__restrict__ out) {
    const float a = 1.0f + seed * 1e-7f, b = seed * 1e-9f + (float)threadIdx.x * 1e-12f;
    float s = 0.f;
    if (MODE == 0) {                                    // 16 x FFMA
        float x[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) x[k] = (float)k;
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int k = 0; k < 16; ++k) x[k] = __fmaf_rn(x[k], a, b);
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) s += x[k];
    } else if (MODE == 1 || MODE == 2 || MODE == 3) {   // 16 x FFMA2 / FMUL2 / FADD2 on 16 independent pairs
        f32x2 x[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) x[k] = pk((float)k, (float)k + 0.5f);
        const f32x2 a2 = bc(a), b2 = bc(b);
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int k = 0; k < 16; ++k) x[k] = MODE == 1 ? fma2(x[k], a2, b2) : (MODE == 2 ? mul2(x[k], a2) : add2(x[k], b2));
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) s += hsum(x[k]);
    } else if (MODE == 4) {                             // 8 x FFMA2 + 8 x FFMA interleaved
        f32x2 x[8]; float y[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { x[k] = pk((float)k, (float)k + 0.5f); y[k] = (float)k; }
        const f32x2 a2 = bc(a), b2 = bc(b);
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int k = 0; k < 8; ++k) { x[k] = fma2(x[k], a2, b2); y[k] = __fmaf_rn(y[k], a, b); }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) s += hsum(x[k]) + y[k];
    } else if (MODE == 5) {                             // 8 x FFMA2 + 8 x FSEL-type (ALU pipe) interleaved
        f32x2 x[8]; float y[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { x[k] = pk((float)k, (float)k + 0.5f); y[k] = (float)k - 3.5f; }
        const f32x2 a2 = bc(a), b2 = bc(b);
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int k = 0; k < 8; ++k) { x[k] = fma2(x[k], a2, b2); y[k] = fsel_asm(y[k], -y[k], b); }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) s += hsum(x[k]) + y[k];
    } else if (MODE == 6) {                             // 16 x (SETP + SELP): the ALU pipe alone
        float y[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) y[k] = (float)k - 7.5f;
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int k = 0; k < 16; ++k) y[k] = fsel_asm(y[k], -y[k], b);
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) s += y[k];
    } else {                                            // MODE 7: 8 x FFMA + 8 x (SETP + SELP)
        float x[8], y[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { x[k] = (float)k; y[k] = (float)k - 3.5f; }
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int k = 0; k < 8; ++k) { x[k] = __fmaf_rn(x[k], a, b); y[k] = fsel_asm(y[k], -y[k], b); }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) s += x[k] + y[k];
    }
    if (s == 123.456f) out[0] = s;                      // never true: keeps the chains alive
}

// *fmas (host) receives the number of fp32 FMAs (modes 0, 1, 4: FFMA2 counts two) or, for the probe modes, of
// instructions of the launch: blocks * 256 threads * iters * 16 [* 2]
int launch_fp32_peak(int blocks, int iters, int mode, float *out, double *fmas, cudaStream_t s) {
    LVDGS_PRE(s);
    switch (mode) {
        case 0: fp32_peak_kernel<0><<<blocks, 256, 0, s>>>(iters, 1.f, out); break;
        case 1: fp32_peak_kernel<1><<<blocks, 256, 0, s>>>(iters, 1.f, out); break;
        case 2: fp32_peak_kernel<2><<<blocks, 256, 0, s>>>(iters, 1.f, out); break;
        case 3: fp32_peak_kernel<3><<<blocks, 256, 0, s>>>(iters, 1.f, out); break;
        case 4: fp32_peak_kernel<4><<<blocks, 256, 0, s>>>(iters, 1.f, out); break;
        case 5: fp32_peak_kernel<5><<<blocks, 256, 0, s>>>(iters, 1.f, out); break;
        case 6: fp32_peak_kernel<6><<<blocks, 256, 0, s>>>(iters, 1.f, out); break;
        default: fp32_peak_kernel<7><<<blocks, 256, 0, s>>>(iters, 1.f, out); break;
    }
    LVDGS_LAUNCHED(s, "fp32_peak");
    const double per_thread = mode == 1 ? 32.0 : (mode == 4 ? 24.0 : 16.0);
    if (fmas) *fmas = (double)blocks * 256.0 * (double)iters * per_thread;
    return 0;
}

}  // namespace lvdgs
