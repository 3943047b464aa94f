// TEST INFRASTRUCTURE: host instantiation of the __host__ __device__ per-face math (gomavatar_b200/csrc/gom_face.cuh)
// so that its hand-derived backward can be checked against float64 autograd of the oracle without a GPU.
// Never linked into libgom_b200.so.
#include "../../gomavatar_b200/csrc/gom_face.cuh"

using namespace gomface;

extern "C" void face_fwd_host(int n, const float *v0, const float *v1, const float *v2, const float *so3,
                              const float *scale, float sigma, float *mean, float *cov6) {
    for (int i = 0; i < n; i++) {
        float R[9], L[9], M[9];
        Frame f;
        so3_exp(so3 + 3 * i, R);
        local_factor(R, scale + 3 * i, L);
        steiner_frame(v0 + 3 * i, v1 + 3 * i, v2 + 3 * i, sigma, mean + 3 * i, f);
        world_cov(f.A, L, M, cov6 + 6 * i);
    }
}

extern "C" void face_bwd_host(int n, const float *v0, const float *v1, const float *v2, const float *so3,
                              const float *scale, float sigma, const float *dmean, const float *dcov6, float *dv0,
                              float *dv1, float *dv2, float *dso3, float *dscale) {
    for (int i = 0; i < n; i++) {
        float R[9], L[9], M[9], mean[3], cov6[6], dL[9], dR[9];
        Frame f;
        so3_exp(so3 + 3 * i, R);
        local_factor(R, scale + 3 * i, L);
        steiner_frame(v0 + 3 * i, v1 + 3 * i, v2 + 3 * i, sigma, mean, f);
        world_cov(f.A, L, M, cov6);
        face_bwd(f, L, M, sigma, dmean + 3 * i, dcov6 + 6 * i, dv0 + 3 * i, dv1 + 3 * i, dv2 + 3 * i, dL);
        local_factor_bwd(R, scale + 3 * i, dL, dR, dscale + 3 * i);
        so3_exp_bwd(so3 + 3 * i, dR, dso3 + 3 * i);
    }
}
