// gom_common.cuh — shared helpers for the sm_100a kernels of libgom_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/gom_b200.h"

// ---------------------------------------------------------------------------------------------------- errors
void gom_set_error(const char *fmt, ...);

#define GOM_REQUIRE(cond, what)                                                     \
    do {                                                                            \
        if (!(cond)) {                                                              \
            gom_set_error("%s: invalid argument: %s", __func__, what);              \
            return GOM_ERR_INVALID;                                                 \
        }                                                                           \
    } while (0)

#define GOM_CUDA(call)                                                              \
    do {                                                                            \
        cudaError_t e_ = (call);                                                    \
        if (e_ != cudaSuccess) {                                                    \
            gom_set_error("%s: %s failed: %s", __func__, #call, cudaGetErrorString(e_)); \
            return GOM_ERR_CUDA;                                                    \
        }                                                                           \
    } while (0)

#define GOM_LAUNCH_CHECK()                                                          \
    do {                                                                            \
        gom_count_launch();                                                         \
        cudaError_t e_ = cudaGetLastError();                                        \
        if (e_ != cudaSuccess) {                                                    \
            gom_set_error("%s: kernel launch failed: %s", __func__, cudaGetErrorString(e_)); \
            return GOM_ERR_CUDA;                                                    \
        }                                                                           \
    } while (0)

// ---------------------------------------------------------------------------- launch counter + per-kernel timers
// gom_launch_count(): kernels this library has launched in this process (bench.py's `gpu_launches` claim).
// gom_profile_*(): optional CUDA-event timers around selected kernels, on the launching stream.
enum GomProfSlot { GOM_PROF_PREPROCESS = 0, GOM_PROF_SCAN, GOM_PROF_EMIT, GOM_PROF_BLEND_FWD, GOM_PROF_BLEND_BWD,
                   GOM_PROF_PREPROCESS_BWD, GOM_PROF_JOINT_FWD, GOM_PROF_JOINT_BWD, GOM_PROF_LBS_FWD, GOM_PROF_LBS_BWD,
                   GOM_PROF_FACE_FWD, GOM_PROF_FACE_BWD, GOM_PROF_PHOTO_FWD, GOM_PROF_PHOTO_BWD, GOM_PROF_CAMERA,
                   GOM_PROF_LPIPS_INPUT, GOM_PROF_BIAS_RELU, GOM_PROF_RELU_BWD, GOM_PROF_LPIPS_TAP_FWD, GOM_PROF_LPIPS_TAP_BWD,
                   GOM_PROF_EVAL_METRICS, GOM_PROF_CONV_FIRST_FWD, GOM_PROF_CONV_FIRST_BWD, GOM_PROF_ADAM,
                   GOM_PROF_MESH_BIN, GOM_PROF_MESH_FWD, GOM_PROF_MESH_BWD,
                   GOM_PROF_SHADOW_COMPACT, GOM_PROF_SHADOW_FWD, GOM_PROF_SHADOW_BWD_DATA,
                   GOM_PROF_SHADOW_BWD_WEIGHTS, GOM_PROF_MESH_REG, GOM_PROF_CONV3X3_FWD, GOM_PROF_CONV3X3_DGRAD,
                   GOM_PROF_WORKLIST, GOM_PROF_TILE_SORT, GOM_PROF_GEMM_TC, GOM_PROF_WGRAD_TC,
                   GOM_PROF_NSLOTS };
void gom_count_launch(void);
void gom_prof_begin(int slot, cudaStream_t stream);
void gom_prof_end(int slot, cudaStream_t stream);

static inline int gom_div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ------------------------------------------------------------------------- programmatic dependent launch (PDL)
// The step is a chain of ~60 (one frame) kernels in one stream / CUDA graph; each boundary costs the drain of the previous
// grid plus the launch and prologue of the next (barrier init, tensor-memory allocation, tensor-map fetch: ~5 us for the
// persistent tcgen05 kernels).  A kernel launched with `gom_launch_pdl` may become resident as soon as every CTA of the
// previous kernel has executed `gom_pdl_trigger()` (or exited) and SM resources are free, runs its prologue, and blocks in
// `gom_pdl_wait()` until the previous grid has completed and its writes are visible.  Rules used throughout: trigger at kernel
// entry; wait before the first global-memory access of any kind other than kernel parameters / weights (reads of produced data
// and ALL writes: an output buffer may be recycled memory the previous kernel still reads).  Launches without the attribute,
// memsets and copies keep full stream order, so a kernel that triggers is safe in front of anything.  Opt-in (GOM_PDL=1): measured
// slower than plain stream order on the LPIPS chain (gom_core.cu), kept for experiments.
#ifdef __CUDACC__
__device__ __forceinline__ void gom_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void gom_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif
bool gom_pdl_enabled(void);
template <typename... KArgs, typename... Args>
static inline cudaError_t gom_launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = gom_pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ------------------------------------------------------------------------------- exactly-rounded fp32 arithmetic
// Every integer decision of the rasterizer (cull, radius, tile rect, depth key) is computed with individually
// rounded fp32 operations in a fixed association order, so that it is bit-identical to the CPU oracle
// (oracle/raster_oracle.c, built with -ffp-contract=off).  nvcc would otherwise contract a*b+c into FMA.
__device__ __forceinline__ float xmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float xadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float xsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float xdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float xsqrt(float a) { return __fsqrt_rn(a); }
// a*b + c*d + e*f  (left-associated)
__device__ __forceinline__ float xdot3(float a, float b, float c, float d, float e, float f) {
    return xadd(xadd(xmul(a, b), xmul(c, d)), xmul(e, f));
}

// App. A.2: first three / four components of the row-vector product [p,1]·M, M = float[16] row-major.
__device__ __forceinline__ void xform4x3(const float *m, const float p[3], float o[3]) {
    o[0] = xadd(xdot3(m[0], p[0], m[4], p[1], m[8], p[2]), m[12]);
    o[1] = xadd(xdot3(m[1], p[0], m[5], p[1], m[9], p[2]), m[13]);
    o[2] = xadd(xdot3(m[2], p[0], m[6], p[1], m[10], p[2]), m[14]);
}
__device__ __forceinline__ float xform_w(const float *m, const float p[3]) {
    return xadd(xdot3(m[3], p[0], m[7], p[1], m[11], p[2]), m[15]);
}

// App. A.3 step 6: upstream's literals are double -> evaluated in fp64, rounded to fp32 on return.
__device__ __forceinline__ float ndc2pix(float v, int S) {
    double d = __dadd_rn((double)v, 1.0);
    d = __dmul_rn(d, (double)S);
    d = __dadd_rn(d, -1.0);
    d = __dmul_rn(d, 0.5);
    return __double2float_rn(d);
}

// Clamped view-space point, rows of J·R and the 2-D covariance (App. A.3 step 3); shared by forward and backward.
struct Cov2D {
    float t[3];
    float xm, ym;          // 0 when the tangent was clamped
    float M0[3], M1[3];    // rows of J·R
    float a, b, c;         // cov2D incl. +0.3 low-pass
    float v0[3], v1[3];    // Sigma·M0^T, Sigma·M1^T
};

__device__ __forceinline__ void cov2d_exact(const float mean[3], const float s[6], const float *view, float fx,
                                            float fy, float tanfovx, float tanfovy, Cov2D &o) {
    float t[3];
    xform4x3(view, mean, t);
    const float limx = xmul(1.3f, tanfovx), limy = xmul(1.3f, tanfovy);
    const float txtz = xdiv(t[0], t[2]), tytz = xdiv(t[1], t[2]);
    o.xm = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
    o.ym = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
    t[0] = xmul(fminf(limx, fmaxf(-limx, txtz)), t[2]);
    t[1] = xmul(fminf(limy, fmaxf(-limy, tytz)), t[2]);
    o.t[0] = t[0]; o.t[1] = t[1]; o.t[2] = t[2];
    const float tz2 = xmul(t[2], t[2]);
    const float J00 = xdiv(fx, t[2]);
    const float J02 = -xdiv(xmul(fx, t[0]), tz2);
    const float J11 = xdiv(fy, t[2]);
    const float J12 = -xdiv(xmul(fy, t[1]), tz2);
#pragma unroll
    for (int k = 0; k < 3; k++) {   // R[r][k] = view[4k + r]
        o.M0[k] = xadd(xmul(J00, view[4 * k + 0]), xmul(J02, view[4 * k + 2]));
        o.M1[k] = xadd(xmul(J11, view[4 * k + 1]), xmul(J12, view[4 * k + 2]));
    }
    const float S[3][3] = {{s[0], s[1], s[2]}, {s[1], s[3], s[4]}, {s[2], s[4], s[5]}};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        o.v0[k] = xdot3(S[k][0], o.M0[0], S[k][1], o.M0[1], S[k][2], o.M0[2]);
        o.v1[k] = xdot3(S[k][0], o.M1[0], S[k][1], o.M1[1], S[k][2], o.M1[2]);
    }
    o.a = xadd(xdot3(o.M0[0], o.v0[0], o.M0[1], o.v0[1], o.M0[2], o.v0[2]), 0.3f);
    o.b = xdot3(o.M0[0], o.v1[0], o.M0[1], o.v1[1], o.M0[2], o.v1[2]);
    o.c = xadd(xdot3(o.M1[0], o.v1[0], o.M1[1], o.v1[1], o.M1[2], o.v1[2]), 0.3f);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// ------------------------------------------------------------------------------------------- per-warp entry culling
// A warp of the blend kernels owns an 8x4-pixel sub-block of a 16x16 tile (lane -> (lane & 7, lane >> 3)).  A list
// entry can only pass the per-pixel test  alpha = min(.99, o * exp(power)) >= 1/255  somewhere in that sub-block if
// the minimum over the sub-block's rectangle of  q = A dx^2 + 2 B dx dy + C dy^2 = -2 power  is at most
// s = 2 ln(255 o).  q is a positive-definite quadratic, so its minimum over a rectangle that does not contain the
// centre lies on the rectangle's boundary: four clamped 1-D minimisations.  The test is conservative (s is inflated by
// 1e-3 relative + 1e-3 against rounding, and the rectangle is the continuous hull of the pixel centres), so skipping
// entries that fail it changes no pixel: they would all have taken the `alpha < 1/255` branch.  Entries are tested
// one per lane (the cost is per 32 entries, not per entry) and the survivors visited through the ballot mask.
__device__ __forceinline__ float q_min_on_vertical(float A, float B, float C, float dx, float ylo, float yhi) {
    const float dy = fminf(fmaxf(__fdividef(-B * dx, C), ylo), yhi);     // argmin over dy of q(dx, dy)
    return A * dx * dx + 2.f * B * dx * dy + C * dy * dy;
}
__device__ __forceinline__ bool entry_reaches_rect(float2 c, float4 co, float rcx, float rcy, float rhw, float rhh) {
    const float ko = 255.0f * co.w;
    const float det = co.x * co.z - co.y * co.y;
    if (!(ko > 1.0f)) return false;                                      // alpha < 1/255 everywhere
    if (!(det > 0.f) || !(co.x > 0.f) || !(co.z > 0.f)) return true;     // degenerate conic: never cull
    const float s = 2.0f * __logf(ko) * 1.001f + 1e-3f;
    // rectangle relative to the centre, d = centre - pixel: dx in [xlo, xhi], dy in [ylo, yhi]
    const float xlo = c.x - rcx - rhw, xhi = c.x - rcx + rhw, ylo = c.y - rcy - rhh, yhi = c.y - rcy + rhh;
    if (xlo <= 0.f && xhi >= 0.f && ylo <= 0.f && yhi >= 0.f) return true;   // centre inside the sub-block
    float q = q_min_on_vertical(co.x, co.y, co.z, xlo, ylo, yhi);
    q = fminf(q, q_min_on_vertical(co.x, co.y, co.z, xhi, ylo, yhi));
    q = fminf(q, q_min_on_vertical(co.z, co.y, co.x, ylo, xlo, xhi));        // horizontal edges: swap the roles of x and y
    q = fminf(q, q_min_on_vertical(co.z, co.y, co.x, yhi, xlo, xhi));
    return q <= s;
}
