// gom_face.cuh — per-face math of the Gaussians-on-Mesh transform: local covariance (so3 exp, scale), Steiner-ellipse
// frame of the posed triangle, world mean / covariance, and the hand-derived backward of all of it.
//
// Reference: models/model.py:27-41 (get_transformation_from_triangle_steiner), :225-234 (centroid, cov_local,
// cov_observation), PyTorch3D 0.7.0 so3_exp_map, models/modules/renderer/gaussian.py:71-75 (upper-
// triangle packing, which is why only the upper triangle of Sigma receives gradient).
//
// Pure functions, __host__ __device__ so that tests/ can compile them for the host and check the derivatives against
// float64 autograd of the oracle without a GPU (that host build is test infrastructure; the product only runs the
// device instantiation inside the kernels of lbs_face.cu).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define GOM_HD __host__ __device__ __forceinline__
#else
#define GOM_HD inline
#endif

namespace gomface {

constexpr float kInv2Sqrt3 = 0.28867513459481287f;   // 1/(2 sqrt 3)
constexpr float kHalfPi = 1.5707963267948966f;

// R = I + (sin th/th) K + ((1-cos th)/th^2) K^2,  th = sqrt(max(|w|^2, 1e-4))   (row-major R[3*r+c])
GOM_HD void so3_exp(const float w[3], float R[9]) {
    const float n2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    const float th = sqrtf(fmaxf(n2, 1e-4f));
    const float inv = 1.0f / th;
    const float f1 = inv * sinf(th), f2 = inv * inv * (1.0f - cosf(th));
    const float x = w[0], y = w[1], z = w[2];
    // K = [[0,-z,y],[z,0,-x],[-y,x,0]];  K^2 = w w^T - |w|^2 I
    R[0] = 1.0f + f2 * (-(y * y) - z * z); R[1] = f1 * -z + f2 * (x * y);        R[2] = f1 * y + f2 * (x * z);
    R[3] = f1 * z + f2 * (x * y);          R[4] = 1.0f + f2 * (-(x * x) - z * z); R[5] = f1 * -x + f2 * (y * z);
    R[6] = f1 * -y + f2 * (x * z);         R[7] = f1 * x + f2 * (y * z);          R[8] = 1.0f + f2 * (-(x * x) - y * y);
}

// dL/dw from dL/dR
GOM_HD void so3_exp_bwd(const float w[3], const float dR[9], float dw[3]) {
    const float x = w[0], y = w[1], z = w[2];
    const float n2 = x * x + y * y + z * z;
    const float th = sqrtf(fmaxf(n2, 1e-4f));
    const float inv = 1.0f / th, s = sinf(th), c = cosf(th);
    const float f1 = inv * s, f2 = inv * inv * (1.0f - c);
    const float K[9] = {0.f, -z, y, z, 0.f, -x, -y, x, 0.f};
    const float K2[9] = {-(y * y) - z * z, x * y, x * z, x * y, -(x * x) - z * z, y * z, x * z, y * z, -(x * x) - y * y};
    float df1 = 0.f, df2 = 0.f;
    for (int i = 0; i < 9; i++) { df1 += dR[i] * K[i]; df2 += dR[i] * K2[i]; }
    // dK = f1 dR + f2 (dR K^T + K^T dR)
    float dK[9];
    for (int r = 0; r < 3; r++)
        for (int cc = 0; cc < 3; cc++) {
            float a = 0.f;
            for (int k = 0; k < 3; k++) a += dR[3 * r + k] * K[3 * cc + k] + K[3 * k + r] * dR[3 * k + cc];
            dK[3 * r + cc] = f1 * dR[3 * r + cc] + f2 * a;
        }
    dw[0] = dK[7] - dK[5];
    dw[1] = dK[2] - dK[6];
    dw[2] = dK[3] - dK[1];
    if (n2 >= 1e-4f) {   // clamp(min=eps) passes gradient only above the threshold
        const float dth = df1 * (c * inv - s * inv * inv) + df2 * (s * inv * inv - 2.0f * (1.0f - c) * inv * inv * inv);
        const float k = dth * inv;
        dw[0] += k * x; dw[1] += k * y; dw[2] += k * z;
    }
}

struct Frame {          // everything the backward needs from the forward of one posed triangle
    float f1[3], f2[3];
    float c0, s0, c1, s1;
    float a0[3], a1[3];
    float n[3], nn;     // a0 x a1 and its norm
    float A[9];         // columns 2 a0 | 2 a1 | sigma n/|n|   (row-major A[3*r+c])
};

GOM_HD void steiner_frame(const float v0[3], const float v1[3], const float v2[3], float sigma, float mean[3], Frame &f) {
    for (int k = 0; k < 3; k++) {
        mean[k] = (v0[k] + v1[k] + v2[k]) / 3.0f;
        f.f1[k] = 0.5f * (v2[k] - mean[k]);
        f.f2[k] = kInv2Sqrt3 * (v1[k] - v0[k]);
    }
    const float p = 2.0f * f.f1[0] * f.f2[0] + 2.0f * f.f1[1] * f.f2[1] + 2.0f * f.f1[2] * f.f2[2];
    const float q = (f.f1[0] * f.f1[0] + f.f1[1] * f.f1[1] + f.f1[2] * f.f1[2]) -
                    (f.f2[0] * f.f2[0] + f.f2[1] * f.f2[1] + f.f2[2] * f.f2[2]);
    const float t0 = atan2f(p, q) * 0.5f;
    f.c0 = cosf(t0); f.s0 = sinf(t0);
    f.c1 = cosf(t0 + kHalfPi); f.s1 = sinf(t0 + kHalfPi);
    for (int k = 0; k < 3; k++) {
        f.a0[k] = f.f1[k] * f.c0 + f.f2[k] * f.s0;
        f.a1[k] = f.f1[k] * f.c1 + f.f2[k] * f.s1;
    }
    f.n[0] = f.a0[1] * f.a1[2] - f.a0[2] * f.a1[1];
    f.n[1] = f.a0[2] * f.a1[0] - f.a0[0] * f.a1[2];
    f.n[2] = f.a0[0] * f.a1[1] - f.a0[1] * f.a1[0];
    f.nn = sqrtf(f.n[0] * f.n[0] + f.n[1] * f.n[1] + f.n[2] * f.n[2]);
    const float sc = sigma / fmaxf(f.nn, 1e-12f);
    for (int r = 0; r < 3; r++) {
        f.A[3 * r + 0] = 2.0f * f.a0[r];
        f.A[3 * r + 1] = 2.0f * f.a1[r];
        f.A[3 * r + 2] = f.n[r] * sc;
    }
}

// L = R diag(s)
GOM_HD void local_factor(const float R[9], const float s[3], float L[9]) {
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) L[3 * r + c] = R[3 * r + c] * s[c];
}

// Sigma = (A L)(A L)^T packed xx,xy,xz,yy,yz,zz; M = A L is returned for the backward
GOM_HD void world_cov(const float A[9], const float L[9], float M[9], float cov6[6]) {
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) M[3 * r + c] = A[3 * r] * L[c] + A[3 * r + 1] * L[3 + c] + A[3 * r + 2] * L[6 + c];
    cov6[0] = M[0] * M[0] + M[1] * M[1] + M[2] * M[2];
    cov6[1] = M[0] * M[3] + M[1] * M[4] + M[2] * M[5];
    cov6[2] = M[0] * M[6] + M[1] * M[7] + M[2] * M[8];
    cov6[3] = M[3] * M[3] + M[4] * M[4] + M[5] * M[5];
    cov6[4] = M[3] * M[6] + M[4] * M[7] + M[5] * M[8];
    cov6[5] = M[6] * M[6] + M[7] * M[7] + M[8] * M[8];
}

// Backward of (mean, cov6) w.r.t. the three posed vertices and the local factor L.
GOM_HD void face_bwd(const Frame &f, const float L[9], const float M[9], float sigma, const float dmean[3],
                     const float g[6], float dv0[3], float dv1[3], float dv2[3], float dL[9]) {
    // dM = Gs M with Gs = G + G^T (only the upper triangle of Sigma is consumed downstream)
    const float Gs[9] = {2.0f * g[0], g[1], g[2], g[1], 2.0f * g[3], g[4], g[2], g[4], 2.0f * g[5]};
    float dM[9], dA[9];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) dM[3 * r + c] = Gs[3 * r] * M[c] + Gs[3 * r + 1] * M[3 + c] + Gs[3 * r + 2] * M[6 + c];
    for (int r = 0; r < 3; r++)
        for (int k = 0; k < 3; k++) {
            dA[3 * r + k] = dM[3 * r] * L[3 * k] + dM[3 * r + 1] * L[3 * k + 1] + dM[3 * r + 2] * L[3 * k + 2];       // dM L^T
            dL[3 * r + k] = f.A[r] * dM[k] + f.A[3 + r] * dM[3 + k] + f.A[6 + r] * dM[6 + k];                      // A^T dM
        }
    float da0[3], da1[3], dnh[3];
    for (int r = 0; r < 3; r++) { da0[r] = 2.0f * dA[3 * r]; da1[r] = 2.0f * dA[3 * r + 1]; dnh[r] = dA[3 * r + 2]; }
    // nh = sigma n / max(|n|, eps)
    float dn[3];
    if (f.nn > 1e-12f) {
        const float inv = 1.0f / f.nn;
        const float nh0 = f.n[0] * inv, nh1 = f.n[1] * inv, nh2 = f.n[2] * inv;
        const float dot = nh0 * dnh[0] + nh1 * dnh[1] + nh2 * dnh[2];
        const float k = sigma * inv;
        dn[0] = k * (dnh[0] - nh0 * dot); dn[1] = k * (dnh[1] - nh1 * dot); dn[2] = k * (dnh[2] - nh2 * dot);
    } else {
        const float k = sigma / 1e-12f;
        dn[0] = k * dnh[0]; dn[1] = k * dnh[1]; dn[2] = k * dnh[2];
    }
    // n = a0 x a1:  da0 += a1 x dn,  da1 += dn x a0
    da0[0] += f.a1[1] * dn[2] - f.a1[2] * dn[1];
    da0[1] += f.a1[2] * dn[0] - f.a1[0] * dn[2];
    da0[2] += f.a1[0] * dn[1] - f.a1[1] * dn[0];
    da1[0] += dn[1] * f.a0[2] - dn[2] * f.a0[1];
    da1[1] += dn[2] * f.a0[0] - dn[0] * f.a0[2];
    da1[2] += dn[0] * f.a0[1] - dn[1] * f.a0[0];
    float df1[3], df2[3], dt0 = 0.f;
    for (int k = 0; k < 3; k++) {
        df1[k] = da0[k] * f.c0 + da1[k] * f.c1;
        df2[k] = da0[k] * f.s0 + da1[k] * f.s1;
        dt0 += da0[k] * (-f.f1[k] * f.s0 + f.f2[k] * f.c0) + da1[k] * (-f.f1[k] * f.s1 + f.f2[k] * f.c1);
    }
    const float p = 2.0f * (f.f1[0] * f.f2[0] + f.f1[1] * f.f2[1] + f.f1[2] * f.f2[2]);
    const float q = (f.f1[0] * f.f1[0] + f.f1[1] * f.f1[1] + f.f1[2] * f.f1[2]) -
                    (f.f2[0] * f.f2[0] + f.f2[1] * f.f2[1] + f.f2[2] * f.f2[2]);
    const float den = p * p + q * q;
    const float dp = 0.5f * dt0 * q / den, dq = -0.5f * dt0 * p / den;   // atan2(0,0): same blow-up as the reference
    for (int k = 0; k < 3; k++) {
        df1[k] += 2.0f * dp * f.f2[k] + 2.0f * dq * f.f1[k];
        df2[k] += 2.0f * dp * f.f1[k] - 2.0f * dq * f.f2[k];
    }
    for (int k = 0; k < 3; k++) {
        const float dc = (dmean[k] - 0.5f * df1[k]) / 3.0f;
        dv0[k] = dc - kInv2Sqrt3 * df2[k];
        dv1[k] = dc + kInv2Sqrt3 * df2[k];
        dv2[k] = dc + 0.5f * df1[k];
    }
}

// dL/dR, dL/ds from dL/dL  (L = R diag(s))
GOM_HD void local_factor_bwd(const float R[9], const float s[3], const float dL[9], float dR[9], float ds[3]) {
    for (int c = 0; c < 3; c++) ds[c] = dL[c] * R[c] + dL[3 + c] * R[3 + c] + dL[6 + c] * R[6 + c];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) dR[3 * r + c] = dL[3 * r + c] * s[c];
}

}  // namespace gomface
