// cub_compare.cu -- NOT on the product path.  Wraps cub::DeviceRadixSort::SortPairs, the library call the
// reference's rasterizer makes for K4, so that bench/tests can time and cross-check the
// hand-written onesweep sort against it on the same box (BASELINE.md B-CUB).
#include "common.cuh"
#include <cub/device/device_radix_sort.cuh>

extern "C" size_t lvdgs_cub_sort_workspace_bytes(int64_t n, int32_t end_bit) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint64_t *)nullptr, (uint64_t *)nullptr,
                                    (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)n, 0, end_bit);
    return bytes;
}

extern "C" int lvdgs_cub_sort_pairs(int64_t n, const uint64_t *keys_in, uint64_t *keys_out, const uint32_t *vals_in,
                                    uint32_t *vals_out, int32_t end_bit, void *workspace, size_t workspace_bytes,
                                    void *stream) {
    cudaError_t e = cub::DeviceRadixSort::SortPairs(workspace, workspace_bytes, keys_in, keys_out, vals_in, vals_out,
                                                    (int)n, 0, end_bit, (cudaStream_t)stream);
    if (e != cudaSuccess) { lvdgs::set_error("cub sort failed: %s", cudaGetErrorString(e)); return 1; }
    return 0;
}
