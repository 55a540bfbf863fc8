/*
 * oracle/knn_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of simple-knn's distCUDA2 (the CUDA source is in the missing
 * submodules.zip -- PARITY UNPINNED): for every point, the mean of the squared distances to its three
 * nearest neighbours (self excluded).  The published algorithm is exact (Morton ordering + conservative
 * box pruning), so the oracle is brute force; squared distances are formed in float32 like the GPU path,
 * the three smallest are averaged as (d1+d2+d3)/3 in float32.
 * Call site that pins the name/shape: GaussianModel.create_pcd_from_image_and_depth (missing file) reached
 * via utils/slam_backend.py:75-78.
 */
#include <stdint.h>
#include <float.h>

void oracle_dist2(int P, const float *pts, float *out) {
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < P; ++i) {
        const float x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
        float b0 = FLT_MAX, b1 = FLT_MAX, b2 = FLT_MAX;
        for (int j = 0; j < P; ++j) {
            if (j == i) continue;
            const float dx = pts[3 * j] - x, dy = pts[3 * j + 1] - y, dz = pts[3 * j + 2] - z;
            const float d = dx * dx + dy * dy + dz * dz;
            if (d < b2) {
                if (d < b1) { b2 = b1; if (d < b0) { b1 = b0; b0 = d; } else b1 = d; }
                else b2 = d;
            }
        }
        out[i] = (b0 + b1 + b2) / 3.f;
    }
}
