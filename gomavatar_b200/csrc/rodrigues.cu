// rodrigues.cu — axis-angle vectors -> rotation matrices and the backward, one launch each (gom_rodrigues_*).
// Reference utils/network_util.py:66-92 (RodriguesModule: theta = sqrt(1e-5 + |r|^2), used by the pose-refinement module
// models/modules/pose_refinement_module.py:39-48, which prepends the identity for the root joint, and by the rigid
// global_R of models/model.py:218-221).  torch evaluates the nine entries with ~50 elementwise kernels on [B * 23] elements
// forward and ~150 backward; at the reference's batch of one frame that is a tenth of the launches of the whole step.
#include <math.h>

#include "gom_common.cuh"

namespace {

constexpr int kThreads = 128;

struct Rod { float x, y, z, c, s, oc, theta; };

__device__ __forceinline__ Rod rod_terms(const float *r, float eps) {
    Rod o;
    o.theta = sqrtf(eps + r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    o.x = r[0] / o.theta; o.y = r[1] / o.theta; o.z = r[2] / o.theta;
    o.c = cosf(o.theta); o.s = sinf(o.theta); o.oc = 1.0f - o.c;
    return o;
}

__global__ void __launch_bounds__(kThreads) k_rodrigues_fwd(GomRodriguesArgs a) {
    const int i = blockIdx.x * kThreads + threadIdx.x;            // output matrix index
    const int per = a.group + a.prepend_identity, total = (a.n_rot / a.group) * per;
    if (i >= total) return;
    float *R = a.R + 9LL * i;
    const int f = i / per, j = i - f * per;
    if (a.prepend_identity && j == 0) {
        R[0] = 1.f; R[1] = 0.f; R[2] = 0.f; R[3] = 0.f; R[4] = 1.f; R[5] = 0.f; R[6] = 0.f; R[7] = 0.f; R[8] = 1.f;
        return;
    }
    const Rod t = rod_terms(a.rvec + 3LL * (f * a.group + j - a.prepend_identity), a.eps);
    const float x = t.x, y = t.y, z = t.z, c = t.c, s = t.s, oc = t.oc;
    R[0] = x * x + (1.0f - x * x) * c; R[1] = x * y * oc - z * s; R[2] = x * z * oc + y * s;
    R[3] = x * y * oc + z * s; R[4] = y * y + (1.0f - y * y) * c; R[5] = y * z * oc - x * s;
    R[6] = x * z * oc - y * s; R[7] = y * z * oc + x * s; R[8] = z * z + (1.0f - z * z) * c;
}

__global__ void __launch_bounds__(kThreads) k_rodrigues_bwd(GomRodriguesArgs a) {
    const int n = blockIdx.x * kThreads + threadIdx.x;            // rotation vector index
    if (n >= a.n_rot) return;
    const int per = a.group + a.prepend_identity, f = n / a.group, j = n - f * a.group;
    const float *G = a.g_R + 9LL * (f * per + j + a.prepend_identity);
    const float *r = a.rvec + 3LL * n;
    const Rod t = rod_terms(r, a.eps);
    const float x = t.x, y = t.y, z = t.z, c = t.c, s = t.s, oc = t.oc;
    // reverse mode through the nine entries
    float gx = G[0] * 2.f * x * oc + G[1] * y * oc + G[2] * z * oc + G[3] * y * oc - G[5] * s + G[6] * z * oc + G[7] * s;
    float gy = G[1] * x * oc + G[2] * s + G[3] * x * oc + G[4] * 2.f * y * oc + G[5] * z * oc - G[6] * s + G[7] * z * oc;
    float gz = -G[1] * s + G[2] * x * oc + G[3] * s + G[5] * y * oc + G[6] * x * oc + G[7] * y * oc + G[8] * 2.f * z * oc;
    const float goc = G[1] * x * y + G[2] * x * z + G[3] * x * y + G[5] * y * z + G[6] * x * z + G[7] * y * z;
    const float gc = G[0] * (1.f - x * x) + G[4] * (1.f - y * y) + G[8] * (1.f - z * z) - goc;
    const float gs = -G[1] * z + G[2] * y + G[3] * z - G[5] * x - G[6] * y + G[7] * x;
    // u = r / theta, theta = sqrt(eps + |r|^2)
    float gtheta = -s * gc + c * gs - (gx * r[0] + gy * r[1] + gz * r[2]) / (t.theta * t.theta);
    float *o = a.g_rvec + 3LL * n;
    o[0] = gx / t.theta + gtheta * r[0] / t.theta;
    o[1] = gy / t.theta + gtheta * r[1] / t.theta;
    o[2] = gz / t.theta + gtheta * r[2] / t.theta;
}

int check(const GomRodriguesArgs *p) {
    GOM_REQUIRE(p != nullptr, "args");
    GOM_REQUIRE(p->n_rot > 0 && p->group > 0 && p->n_rot % p->group == 0, "n_rot must be a positive multiple of group");
    GOM_REQUIRE(p->prepend_identity == 0 || p->prepend_identity == 1, "prepend_identity");
    GOM_REQUIRE(p->rvec, "null pointer");
    return GOM_OK;
}

}  // namespace

extern "C" int gom_rodrigues_forward(const GomRodriguesArgs *p, gom_stream_t stream) {
    if (int rc = check(p)) return rc;
    GOM_REQUIRE(p->R, "null output");
    const int total = (p->n_rot / p->group) * (p->group + p->prepend_identity);
    k_rodrigues_fwd<<<gom_div_up(total, kThreads), kThreads, 0, (cudaStream_t)stream>>>(*p);
    GOM_LAUNCH_CHECK();
    return GOM_OK;
}

extern "C" int gom_rodrigues_backward(const GomRodriguesArgs *p, gom_stream_t stream) {
    if (int rc = check(p)) return rc;
    GOM_REQUIRE(p->g_R && p->g_rvec, "null gradient pointer");
    k_rodrigues_bwd<<<gom_div_up(p->n_rot, kThreads), kThreads, 0, (cudaStream_t)stream>>>(*p);
    GOM_LAUNCH_CHECK();
    return GOM_OK;
}

extern "C" size_t gom_sizeof_rodrigues_args(void) { return sizeof(GomRodriguesArgs); }
