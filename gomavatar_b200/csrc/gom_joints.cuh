// gom_joints.cuh — small affine-matrix helpers for the skeleton chain (reference utils/body_util.py:591-638).
// __host__ __device__ so tests can instantiate them on the host (test infrastructure only).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define GOM_HD __host__ __device__ __forceinline__
#else
#define GOM_HD inline
#endif

namespace gomjoint {

// General 4x4 inverse (Gauss-Jordan, partial pivoting) — the reference calls torch.inverse on cnl_gtfms, which in
// its datasets are pure translations, but the interface accepts any invertible matrix.  Row-major.
GOM_HD bool inverse4x4(const float m[16], float inv[16]) {
    float a[4][8];
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) {
            a[r][c] = m[4 * r + c];
            a[r][4 + c] = (r == c) ? 1.0f : 0.0f;
        }
    bool ok = true;
    for (int col = 0; col < 4; col++) {
        int piv = col;
        float best = fabsf(a[col][col]);
        for (int r = col + 1; r < 4; r++) {
            const float v = fabsf(a[r][col]);
            if (v > best) { best = v; piv = r; }
        }
        if (best == 0.0f) ok = false;
        if (piv != col)
            for (int c = 0; c < 8; c++) { const float t = a[col][c]; a[col][c] = a[piv][c]; a[piv][c] = t; }
        const float ip = 1.0f / a[col][col];
        for (int c = 0; c < 8; c++) a[col][c] *= ip;
        for (int r = 0; r < 4; r++) {
            if (r == col) continue;
            const float f = a[r][col];
            if (f != 0.0f)
                for (int c = 0; c < 8; c++) a[r][c] -= f * a[col][c];
        }
    }
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) inv[4 * r + c] = a[r][4 + c];
    return ok;
}

// C = A·B for 3x4 affine blocks with implicit bottom row [0 0 0 1]  (row-major, 12 floats each)
GOM_HD void affine_mul(const float A[12], const float B[12], float C[12]) {
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 4; c++) {
            float s = A[4 * r] * B[c] + A[4 * r + 1] * B[4 + c] + A[4 * r + 2] * B[8 + c];
            if (c == 3) s += A[4 * r + 3];
            C[4 * r + c] = s;
        }
}

}  // namespace gomjoint
