// Is MUFU.RCP exact at 1.0 (and at other powers of two)?  The branch-free blend backward multiplies T by rcp(1 - 0).
#include <cstdio>
__global__ void k(float *out) {
    float xs[6] = {1.0f, 2.0f, 0.5f, 4.0f, 0.25f, 1.0f - 0.0f};
    for (int i = 0; i < 6; ++i) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(xs[i])); out[i] = y; }
}
int main() {
    float *d, h[6];
    cudaMalloc(&d, sizeof h); k<<<1, 1>>>(d); cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    for (int i = 0; i < 6; ++i) printf("%.9g ", h[i]);
    printf("\n");
    return 0;
}
