// lpips.cu — the non-GEMM half of LPIPS-VGG (v0.1), forward and backward, fused (sm_100a).
//
// Replaces, for reference utils/lpips/lpips.py:81-123 (+ ScalingLayer :126-133, NetLinLayer :136-146,
// normalize_tensor __init__.py:40-42, the ReLU / MaxPool2d layers of pretrained_networks.py:96-134) as called from
// train.py:113-121, everything that is NOT a convolution: input scaling, bias + ReLU, 2x2 max-pooling, channel-unit
// normalisation, squared difference, the 1x1 "lin" heads, the spatial mean, and ALL of their backward passes
// (autograd in the reference: ~25 elementwise/reduction launches per tapped layer, each a full pass over HBM).
// The 3x3 convolutions themselves stay library tensor-core GEMMs (cuDNN), see gomavatar_b200/lpips.py.
//
// Layout: activations are NHWC fp32 ("channels_last"), so one pixel's channel vector is contiguous.  A batch holds
// the B predicted images first and their B targets after them ([2B,h,w,C]); a warp owns one 2x2 pixel quad of one
// (prediction, target) pair, lanes stride over channels with 8/16-byte loads, channel sums are warp shuffles.
// Every kernel is a single pass over its tensors: HBM-bound by construction (roofline in DESIGN.md §4).
#include "gom_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr float kEps = 1e-10f;          // normalize_tensor eps (inside and outside the sqrt)

// ------------------------------------------------------------------------------------------------ input scaling
// x = (v * mul + add - shift_c) / scale_c ; mul/add = (2,-1) when the caller hands [0,1] images (train.py:114-116).
__constant__ float c_shift[3] = {-.030f, -.088f, -.188f};
__constant__ float c_scale[3] = {.458f, .448f, .450f};

// A thread handles 12 consecutive floats = 4 pixels (three float4): the channel of every element is then a compile-time
// constant (no 64-bit modulo per element) and all accesses are 16-byte vectors.  VEC = false: any size / alignment.
template <bool VEC>
__global__ void __launch_bounds__(kThreads) k_lpips_input_fwd(const float *pred, const float *gt, float *out,
                                                              long long n_half, float mul, float add) {
    gom_pdl_trigger();            // programmatic dependent launch (gom_common.cuh)
    gom_pdl_wait();
    // n_half = B*H*W*3 elements per half; out = [pred half | gt half]
    if (VEC) {
        const long long groups_half = n_half / 12;
        for (long long gi = (long long)blockIdx.x * kThreads + threadIdx.x; gi < 2 * groups_half; gi += (long long)gridDim.x * kThreads) {
            const bool second = gi >= groups_half;
            const float4 *src = reinterpret_cast<const float4 *>(second ? gt : pred) + (second ? gi - groups_half : gi) * 3;
            float4 *dst = reinterpret_cast<float4 *>(out) + gi * 3;
            float v[12];
#pragma unroll
            for (int q = 0; q < 3; q++) { const float4 f = __ldg(src + q); v[4 * q] = f.x; v[4 * q + 1] = f.y; v[4 * q + 2] = f.z; v[4 * q + 3] = f.w; }
#pragma unroll
            for (int j = 0; j < 12; j++) v[j] = (v[j] * mul + add - c_shift[j % 3]) / c_scale[j % 3];
#pragma unroll
            for (int q = 0; q < 3; q++) dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
        return;
    }
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < 2 * n_half; i += (long long)gridDim.x * kThreads) {
        const float v = i < n_half ? pred[i] : gt[i - n_half];
        const int c = (int)(i % 3);
        out[i] = (v * mul + add - c_shift[c]) / c_scale[c];
    }
}

template <bool VEC>
__global__ void __launch_bounds__(kThreads) k_lpips_input_bwd(const float *g, float *d_pred, long long n_half, float mul) {
    gom_pdl_trigger();
    gom_pdl_wait();
    if (VEC) {
        for (long long gi = (long long)blockIdx.x * kThreads + threadIdx.x; gi < n_half / 12; gi += (long long)gridDim.x * kThreads) {
            const float4 *src = reinterpret_cast<const float4 *>(g) + gi * 3;
            float4 *dst = reinterpret_cast<float4 *>(d_pred) + gi * 3;
            float v[12];
#pragma unroll
            for (int q = 0; q < 3; q++) { const float4 f = __ldg(src + q); v[4 * q] = f.x; v[4 * q + 1] = f.y; v[4 * q + 2] = f.z; v[4 * q + 3] = f.w; }
#pragma unroll
            for (int j = 0; j < 12; j++) v[j] = v[j] * mul / c_scale[j % 3];
#pragma unroll
            for (int q = 0; q < 3; q++) dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
        return;
    }
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n_half; i += (long long)gridDim.x * kThreads)
        d_pred[i] = g[i] * mul / c_scale[(int)(i % 3)];
}

// ------------------------------------------------------------------------------------------------ bias + ReLU
// in place over [n_pix, C], C % 4 == 0.
__global__ void __launch_bounds__(kThreads) k_bias_relu(float4 *x, const float4 *bias, long long n4, int c4) {
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n4; i += (long long)gridDim.x * kThreads) {
        float4 v = x[i];
        const float4 b = __ldg(bias + (int)(i % c4));
        v.x = fmaxf(v.x + b.x, 0.f); v.y = fmaxf(v.y + b.y, 0.f); v.z = fmaxf(v.z + b.z, 0.f); v.w = fmaxf(v.w + b.w, 0.f);
        x[i] = v;
    }
}

// grad *= (act > 0), in place
__global__ void __launch_bounds__(kThreads) k_relu_bwd(const float4 *act, float4 *grad, long long n4) {
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n4; i += (long long)gridDim.x * kThreads) {
        const float4 a = act[i];
        float4 g = grad[i];
        g.x = a.x > 0.f ? g.x : 0.f; g.y = a.y > 0.f ? g.y : 0.f; g.z = a.z > 0.f ? g.z : 0.f; g.w = a.w > 0.f ? g.w : 0.f;
        grad[i] = g;
    }
}

// ------------------------------------------------------------------------------------------------ tapped layers
struct TapDev {
    int B, h, w, pool;
    const float *feats;         // [2B,h,w,C]
    const float *lin;           // [C]
    float *layer_sums;          // [B]   += spatial mean of sum_c lin_c (u0_c - u1_c)^2
    float *pooled;              // [2B,h/2,w/2,C]
    const float *dval;          // [B]
    const float *d_pooled;      // [B,h/2,w/2,C]
    float *d_pre;               // [B,h,w,C]
};

// Lane layout: a pixel's C channels are spread over LPP = min(32, C/4) lanes with 16-byte loads, so a warp holds
// PPW = 32/LPP pixels at once (C = 64: two pixels per warp, the quad in NIT = 2 iterations) and channel sums are
// xor-shuffles inside the LPP-lane group.  Register r of a lane is channel (r/4)*4*LPP + 4*sl + r%4, sl = lane % LPP.
template <int C> struct TL {
    static_assert(C % 32 == 0 && C >= 32, "channels must be a multiple of 32");
    static constexpr int LPP = C / 4 < 32 ? C / 4 : 32;  // lanes per pixel
    static constexpr int PPW = 32 / LPP;                 // pixels per warp (1, 2 or 4)
    static constexpr int CPL = C / LPP;                  // channels per lane (4, 8 or 16)
    static constexpr int NV = CPL / 4;                   // float4 loads per pixel per lane
    static constexpr int NIT = 4 / PPW;                  // iterations to cover a 2x2 quad
};

template <int C> __device__ __forceinline__ void load_px(const float *p, int sl, float (&f)[TL<C>::CPL]) {
    using L = TL<C>;
#pragma unroll
    for (int k = 0; k < L::NV; k++) {
        const float4 v = *reinterpret_cast<const float4 *>(p + k * 4 * L::LPP + 4 * sl);
        f[4 * k] = v.x; f[4 * k + 1] = v.y; f[4 * k + 2] = v.z; f[4 * k + 3] = v.w;
    }
}

template <int C> __device__ __forceinline__ void store_px(float *p, int sl, const float (&f)[TL<C>::CPL]) {
    using L = TL<C>;
#pragma unroll
    for (int k = 0; k < L::NV; k++)
        *reinterpret_cast<float4 *>(p + k * 4 * L::LPP + 4 * sl) = make_float4(f[4 * k], f[4 * k + 1], f[4 * k + 2], f[4 * k + 3]);
}

template <int C> __device__ __forceinline__ void load_lin(const float *lin, int sl, float (&f)[TL<C>::CPL]) {
    using L = TL<C>;
#pragma unroll
    for (int r = 0; r < L::CPL; r++) f[r] = __ldg(lin + (r / 4) * 4 * L::LPP + 4 * sl + (r % 4));
}

// 1 / (sqrt(s + eps) + eps) and sqrt(s + eps) with the SFU approximations (MUFU.RSQ / MUFU.RCP, <= 2 ulp each): the
// IEEE sqrt + divide sequences were ~40 of the ~110 instructions per pixel pair of an otherwise HBM-bound kernel.
__device__ __forceinline__ float inv_norm(float s, float &n) {
    const float se = s + kEps;                 // >= 1e-10: a normal float, rsqrt is finite
    n = se * rsqrtf(se);
    return __fdividef(1.f, n + kEps);
}

template <int LPP> __device__ __forceinline__ void group_sum2(float &a, float &b) {
#pragma unroll
    for (int d = LPP / 2; d > 0; d >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, d);
        b += __shfl_xor_sync(0xffffffffu, b, d);
    }
}
template <int LPP> __device__ __forceinline__ float group_sum(float a) {
#pragma unroll
    for (int d = LPP / 2; d > 0; d >>= 1) a += __shfl_xor_sync(0xffffffffu, a, d);
    return a;
}

// Forward: per (prediction, target) pair and pixel  d = sum_c lin_c (f0_c/(n0+eps) - f1_c/(n1+eps))^2,
// n = sqrt(sum_c f_c^2 + eps); layer_sums[b] += mean over pixels; optionally the 2x2/2 max-pool of BOTH halves.
template <int C>
__global__ void __launch_bounds__(kThreads) k_lpips_tap_fwd(TapDev a) {
    using L = TL<C>;
    gom_pdl_trigger();
    gom_pdl_wait();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int grp = lane / L::LPP, sl = lane % L::LPP;
    const int b = blockIdx.y;
    const int qw = (a.w + 1) >> 1, qh = (a.h + 1) >> 1, nq = qw * qh;
    const int ph = a.h >> 1, pw = a.w >> 1;
    const long long img = (long long)a.h * a.w * C;
    const float *F0 = a.feats + (long long)b * img, *F1 = a.feats + (long long)(b + a.B) * img;
    float lin[L::CPL];
    load_lin<C>(a.lin, sl, lin);
    float acc = 0.f;
    const int q0 = blockIdx.x * kWarps + wid, qstride = gridDim.x * kWarps;
    const int dqy = qstride / qw, dqx = qstride - dqy * qw;
    int qy = q0 / qw, qx = q0 - qy * qw;
    // For the wide-image levels (C <= 128: 16 registers per pixel pair) the NEXT quad's features are requested before the
    // current quad is reduced, so that every warp always has a full set of loads in flight (the kernel is HBM-bound and
    // otherwise alternates between a load phase and a shuffle/reduce phase).
    constexpr bool kPrefetch = (L::CPL * L::NIT <= 16);
    float c0[kPrefetch ? L::NIT : 1][L::CPL], c1[kPrefetch ? L::NIT : 1][L::CPL];
    auto load_quad = [&](int qy_, int qx_, float (&g0)[kPrefetch ? L::NIT : 1][L::CPL], float (&g1)[kPrefetch ? L::NIT : 1][L::CPL]) {
#pragma unroll
        for (int it = 0; it < (kPrefetch ? L::NIT : 1); it++) {
            const int s_ = it * L::PPW + grp;
            const int y = 2 * qy_ + (s_ >> 1), x = 2 * qx_ + (s_ & 1);
            if (y < a.h && x < a.w) {
                const long long o = ((long long)y * a.w + x) * C;
                load_px<C>(F0 + o, sl, g0[it]);
                load_px<C>(F1 + o, sl, g1[it]);
            } else {
#pragma unroll
                for (int r = 0; r < L::CPL; r++) { g0[it][r] = 0.f; g1[it][r] = 0.f; }
            }
        }
    };
    if (kPrefetch && q0 < nq) load_quad(qy, qx, c0, c1);
    for (int q = q0; q < nq; q += qstride, qy += dqy, qx += dqx) {
        if (qx >= qw) { qx -= qw; qy++; }
        float n0_[kPrefetch ? L::NIT : 1][L::CPL], n1_[kPrefetch ? L::NIT : 1][L::CPL];
        if (kPrefetch && q + qstride < nq) {
            int nqy = qy + dqy, nqx = qx + dqx;
            if (nqx >= qw) { nqx -= qw; nqy++; }
            load_quad(nqy, nqx, n0_, n1_);
        }
        float m0[L::CPL], m1[L::CPL];
#pragma unroll
        for (int r = 0; r < L::CPL; r++) { m0[r] = -INFINITY; m1[r] = -INFINITY; }
#pragma unroll
        for (int it = 0; it < L::NIT; it++) {
            const int s = it * L::PPW + grp;
            const int y = 2 * qy + (s >> 1), x = 2 * qx + (s & 1);
            const bool ok = y < a.h && x < a.w;
            const long long o = ((long long)y * a.w + x) * C;
            float f0[L::CPL], f1[L::CPL];
            if (kPrefetch) {
#pragma unroll
                for (int r = 0; r < L::CPL; r++) { f0[r] = c0[kPrefetch ? it : 0][r]; f1[r] = c1[kPrefetch ? it : 0][r]; }
            } else if (ok) {
                load_px<C>(F0 + o, sl, f0);
                load_px<C>(F1 + o, sl, f1);
            } else {
#pragma unroll
                for (int r = 0; r < L::CPL; r++) { f0[r] = 0.f; f1[r] = 0.f; }
            }
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int r = 0; r < L::CPL; r++) {
                s0 += f0[r] * f0[r]; s1 += f1[r] * f1[r];
                if (ok) { m0[r] = fmaxf(m0[r], f0[r]); m1[r] = fmaxf(m1[r], f1[r]); }
            }
            group_sum2<L::LPP>(s0, s1);
            float n0, n1;
            const float i0 = inv_norm(s0, n0), i1 = inv_norm(s1, n1);
            float d = 0.f;
#pragma unroll
            for (int r = 0; r < L::CPL; r++) {
                const float t = f0[r] * i0 - f1[r] * i1;
                d += lin[r] * t * t;
            }
            if (ok) acc += d;                                     // reduced across lanes once, at the end
        }
        if (a.pool && qy < ph && qx < pw) {                       // warp-uniform; all four pixels exist
#pragma unroll
            for (int dlt = L::LPP; dlt < 32; dlt <<= 1) {
#pragma unroll
                for (int r = 0; r < L::CPL; r++) {
                    m0[r] = fmaxf(m0[r], __shfl_xor_sync(0xffffffffu, m0[r], dlt));
                    m1[r] = fmaxf(m1[r], __shfl_xor_sync(0xffffffffu, m1[r], dlt));
                }
            }
            if (grp == 0) {
                const long long po = ((long long)qy * pw + qx) * C, pimg = (long long)ph * pw * C;
                store_px<C>(a.pooled + (long long)b * pimg + po, sl, m0);
                store_px<C>(a.pooled + (long long)(b + a.B) * pimg + po, sl, m1);
            }
        }
        if (kPrefetch) {
#pragma unroll
            for (int it = 0; it < L::NIT; it++)
#pragma unroll
                for (int r = 0; r < L::CPL; r++) { c0[kPrefetch ? it : 0][r] = n0_[kPrefetch ? it : 0][r]; c1[kPrefetch ? it : 0][r] = n1_[kPrefetch ? it : 0][r]; }
        }
    }
    __shared__ float sh[kWarps];
    acc = warp_sum(acc);
    if (lane == 0) sh[wid] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < kWarps; k++) t += sh[k];
        atomicAdd(a.layer_sums + b, t / (float)((long long)a.h * a.w));
    }
}

// Backward: gradient wrt the PRE-ReLU convolution output of the prediction half:
//   [ dval_b * d(mean d)/df0  +  max-pool backward of d_pooled (first maximum in scan order, as torch) ] * (f0 > 0)
template <int C>
__global__ void __launch_bounds__(kThreads) k_lpips_tap_bwd(TapDev a) {
    using L = TL<C>;
    gom_pdl_trigger();
    gom_pdl_wait();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int grp = lane / L::LPP, sl = lane % L::LPP;
    const int b = blockIdx.y;
    const int qw = (a.w + 1) >> 1, qh = (a.h + 1) >> 1, nq = qw * qh;
    const int ph = a.h >> 1, pw = a.w >> 1;
    const long long img = (long long)a.h * a.w * C;
    const float *F0 = a.feats + (long long)b * img, *F1 = a.feats + (long long)(b + a.B) * img;
    float *G = a.d_pre + (long long)b * img;
    const float kscale = 2.f * a.dval[b] / (float)((long long)a.h * a.w);
    float lin[L::CPL];
    load_lin<C>(a.lin, sl, lin);
    const int q0 = blockIdx.x * kWarps + wid, qstride = gridDim.x * kWarps;
    const int dqy = qstride / qw, dqx = qstride - dqy * qw;
    int qy = q0 / qw, qx = q0 - qy * qw;
    for (int q = q0; q < nq; q += qstride, qy += dqy, qx += dqx) {
        if (qx >= qw) { qx -= qw; qy++; }
        const bool pooled = a.pool && a.d_pooled && qy < ph && qx < pw;      // warp-uniform; quad complete
        float f0[L::NIT][L::CPL];
        bool ok[L::NIT];
#pragma unroll
        for (int it = 0; it < L::NIT; it++) {
            const int s = it * L::PPW + grp;
            const int y = 2 * qy + (s >> 1), x = 2 * qx + (s & 1);
            ok[it] = y < a.h && x < a.w;
            if (ok[it]) load_px<C>(F0 + ((long long)y * a.w + x) * C, sl, f0[it]);
            else {
#pragma unroll
                for (int r = 0; r < L::CPL; r++) f0[it][r] = 0.f;
            }
        }
        // which of the quad's four pixels (scan order s = 0..3) holds the FIRST maximum of each channel
        float gp[L::CPL];
        uint32_t am = 0;                                                     // 2 bits per register
        if (pooled) {
            load_px<C>(a.d_pooled + ((long long)b * ph * pw + (long long)qy * pw + qx) * C, sl, gp);
#pragma unroll
            for (int r = 0; r < L::CPL; r++) {
                float v4[4];                                                 // the quad's four values of this channel
                if constexpr (L::PPW == 1) {
#pragma unroll
                    for (int s = 0; s < 4; s++) v4[s] = f0[s][r];
                } else if constexpr (L::PPW == 2) {                          // s = 2 it + group: own or the partner group's
#pragma unroll
                    for (int it = 0; it < 2; it++) {
                        const float own = f0[it][r], oth = __shfl_xor_sync(0xffffffffu, own, L::LPP);
                        v4[2 * it] = grp ? oth : own;                        // group 0's pixel comes first in scan order
                        v4[2 * it + 1] = grp ? own : oth;
                    }
                } else {
#pragma unroll
                    for (int s = 0; s < 4; s++) v4[s] = __shfl_sync(0xffffffffu, f0[0][r], s * L::LPP + sl);
                }
                float best = v4[0];
                uint32_t arg = 0;
#pragma unroll
                for (int s = 1; s < 4; s++)
                    if (v4[s] > best || v4[s] != v4[s]) { best = v4[s]; arg = s; }   // strict: the first maximum wins
                am |= arg << (2 * r);
            }
        }
#pragma unroll
        for (int it = 0; it < L::NIT; it++) {
            const int s = it * L::PPW + grp;
            const int y = 2 * qy + (s >> 1), x = 2 * qx + (s & 1);
            const long long o = ((long long)y * a.w + x) * C;
            float f1[L::CPL];
            if (ok[it]) load_px<C>(F1 + o, sl, f1);
            else {
#pragma unroll
                for (int r = 0; r < L::CPL; r++) f1[r] = 0.f;
            }
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int r = 0; r < L::CPL; r++) { s0 += f0[it][r] * f0[it][r]; s1 += f1[r] * f1[r]; }
            group_sum2<L::LPP>(s0, s1);
            float n0, n1;
            const float i0 = inv_norm(s0, n0), i1 = inv_norm(s1, n1);
            float e[L::CPL], dot = 0.f;
#pragma unroll
            for (int r = 0; r < L::CPL; r++) {
                e[r] = kscale * lin[r] * (f0[it][r] * i0 - f1[r] * i1);
                dot += e[r] * f0[it][r];
            }
            dot = group_sum<L::LPP>(dot);
            const float k2 = __fdividef(dot * i0 * i0, n0);
            float g[L::CPL];
#pragma unroll
            for (int r = 0; r < L::CPL; r++) {
                float v = e[r] * i0 - k2 * f0[it][r];
                if (pooled && ((am >> (2 * r)) & 3u) == (uint32_t)s) v += gp[r];
                g[r] = f0[it][r] > 0.f ? v : 0.f;
            }
            if (ok[it]) store_px<C>(G + o, sl, g);
        }
    }
}

// The same kernel for the wide-image levels (C <= 128), with the NEXT quad's operands (both feature halves and the pooled
// gradient) requested before the current quad is processed — see k_lpips_tap_fwd.
// Backward: gradient wrt the PRE-ReLU convolution output of the prediction half:
//   [ dval_b * d(mean d)/df0  +  max-pool backward of d_pooled (first maximum in scan order, as torch) ] * (f0 > 0)
template <int C>
__global__ void __launch_bounds__(kThreads) k_lpips_tap_bwd_pf(TapDev a) {
    using L = TL<C>;
    gom_pdl_trigger();
    gom_pdl_wait();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int grp = lane / L::LPP, sl = lane % L::LPP;
    const int b = blockIdx.y;
    const int qw = (a.w + 1) >> 1, qh = (a.h + 1) >> 1, nq = qw * qh;
    const int ph = a.h >> 1, pw = a.w >> 1;
    const long long img = (long long)a.h * a.w * C;
    const float *F0 = a.feats + (long long)b * img, *F1 = a.feats + (long long)(b + a.B) * img;
    float *G = a.d_pre + (long long)b * img;
    const float kscale = 2.f * a.dval[b] / (float)((long long)a.h * a.w);
    float lin[L::CPL];
    load_lin<C>(a.lin, sl, lin);
    const int q0 = blockIdx.x * kWarps + wid, qstride = gridDim.x * kWarps;
    const int dqy = qstride / qw, dqx = qstride - dqy * qw;
    int qy = q0 / qw, qx = q0 - qy * qw;
    float c0[L::NIT][L::CPL], c1[L::NIT][L::CPL], cgp[L::CPL];
    auto load_quad = [&](int qy_, int qx_, float (&g0)[L::NIT][L::CPL], float (&g1)[L::NIT][L::CPL], float (&ggp)[L::CPL]) {
#pragma unroll
        for (int it = 0; it < L::NIT; it++) {
            const int s_ = it * L::PPW + grp;
            const int y = 2 * qy_ + (s_ >> 1), x = 2 * qx_ + (s_ & 1);
            if (y < a.h && x < a.w) {
                const long long o = ((long long)y * a.w + x) * C;
                load_px<C>(F0 + o, sl, g0[it]);
                load_px<C>(F1 + o, sl, g1[it]);
            } else {
#pragma unroll
                for (int r = 0; r < L::CPL; r++) { g0[it][r] = 0.f; g1[it][r] = 0.f; }
            }
        }
        if (a.pool && a.d_pooled && qy_ < ph && qx_ < pw)
            load_px<C>(a.d_pooled + ((long long)b * ph * pw + (long long)qy_ * pw + qx_) * C, sl, ggp);
    };
    if (q0 < nq) load_quad(qy, qx, c0, c1, cgp);
    for (int q = q0; q < nq; q += qstride, qy += dqy, qx += dqx) {
        if (qx >= qw) { qx -= qw; qy++; }
        const bool pooled = a.pool && a.d_pooled && qy < ph && qx < pw;      // warp-uniform; quad complete
        float n0_[L::NIT][L::CPL], n1_[L::NIT][L::CPL], ngp[L::CPL];
        if (q + qstride < nq) {
            int nqy = qy + dqy, nqx = qx + dqx;
            if (nqx >= qw) { nqx -= qw; nqy++; }
            load_quad(nqy, nqx, n0_, n1_, ngp);
        }
        float f0[L::NIT][L::CPL];
        bool ok[L::NIT];
#pragma unroll
        for (int it = 0; it < L::NIT; it++) {
            const int s = it * L::PPW + grp;
            const int y = 2 * qy + (s >> 1), x = 2 * qx + (s & 1);
            ok[it] = y < a.h && x < a.w;
#pragma unroll
            for (int r = 0; r < L::CPL; r++) f0[it][r] = c0[it][r];
        }
        // which of the quad's four pixels (scan order s = 0..3) holds the FIRST maximum of each channel
        float gp[L::CPL];
        uint32_t am = 0;                                                     // 2 bits per register
        if (pooled) {
#pragma unroll
            for (int r = 0; r < L::CPL; r++) gp[r] = cgp[r];
#pragma unroll
            for (int r = 0; r < L::CPL; r++) {
                float v4[4];                                                 // the quad's four values of this channel
                if constexpr (L::PPW == 1) {
#pragma unroll
                    for (int s = 0; s < 4; s++) v4[s] = f0[s][r];
                } else if constexpr (L::PPW == 2) {                          // s = 2 it + group: own or the partner group's
#pragma unroll
                    for (int it = 0; it < 2; it++) {
                        const float own = f0[it][r], oth = __shfl_xor_sync(0xffffffffu, own, L::LPP);
                        v4[2 * it] = grp ? oth : own;                        // group 0's pixel comes first in scan order
                        v4[2 * it + 1] = grp ? own : oth;
                    }
                } else {
#pragma unroll
                    for (int s = 0; s < 4; s++) v4[s] = __shfl_sync(0xffffffffu, f0[0][r], s * L::LPP + sl);
                }
                float best = v4[0];
                uint32_t arg = 0;
#pragma unroll
                for (int s = 1; s < 4; s++)
                    if (v4[s] > best || v4[s] != v4[s]) { best = v4[s]; arg = s; }   // strict: the first maximum wins
                am |= arg << (2 * r);
            }
        }
#pragma unroll
        for (int it = 0; it < L::NIT; it++) {
            const int s = it * L::PPW + grp;
            const int y = 2 * qy + (s >> 1), x = 2 * qx + (s & 1);
            const long long o = ((long long)y * a.w + x) * C;
            float f1[L::CPL];
#pragma unroll
            for (int r = 0; r < L::CPL; r++) f1[r] = c1[it][r];
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int r = 0; r < L::CPL; r++) { s0 += f0[it][r] * f0[it][r]; s1 += f1[r] * f1[r]; }
            group_sum2<L::LPP>(s0, s1);
            float n0, n1;
            const float i0 = inv_norm(s0, n0), i1 = inv_norm(s1, n1);
            float e[L::CPL], dot = 0.f;
#pragma unroll
            for (int r = 0; r < L::CPL; r++) {
                e[r] = kscale * lin[r] * (f0[it][r] * i0 - f1[r] * i1);
                dot += e[r] * f0[it][r];
            }
            dot = group_sum<L::LPP>(dot);
            const float k2 = __fdividef(dot * i0 * i0, n0);
            float g[L::CPL];
#pragma unroll
            for (int r = 0; r < L::CPL; r++) {
                float v = e[r] * i0 - k2 * f0[it][r];
                if (pooled && ((am >> (2 * r)) & 3u) == (uint32_t)s) v += gp[r];
                g[r] = f0[it][r] > 0.f ? v : 0.f;
            }
            if (ok[it]) store_px<C>(G + o, sl, g);
        }
#pragma unroll
        for (int it = 0; it < L::NIT; it++)
#pragma unroll
            for (int r = 0; r < L::CPL; r++) { c0[it][r] = n0_[it][r]; c1[it][r] = n1_[it][r]; }
#pragma unroll
        for (int r = 0; r < L::CPL; r++) cgp[r] = ngp[r];
    }
}

int grid_for(long long work_items, int per_block) {
    long long g = (work_items + per_block - 1) / per_block;
    const long long cap = 148LL * 16;            // 148 SMs x resident blocks; grid-stride loops cover the rest
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

template <int C> int launch_tap(const TapDev &d, bool bwd, cudaStream_t stream) {
    const int nq = ((d.w + 1) / 2) * ((d.h + 1) / 2);
    int gx = (nq + kWarps - 1) / kWarps;
    const int cap = (148 * 8 + d.B - 1) / d.B;   // ~8 resident blocks per SM over all images
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    dim3 grid(gx, d.B);
    if (bwd) {
        if constexpr (TL<C>::CPL * TL<C>::NIT <= 16) GOM_CUDA(gom_launch_pdl(k_lpips_tap_bwd_pf<C>, grid, dim3(kThreads), 0, stream, d));
        else GOM_CUDA(gom_launch_pdl(k_lpips_tap_bwd<C>, grid, dim3(kThreads), 0, stream, d));
    }
    else GOM_CUDA(gom_launch_pdl(k_lpips_tap_fwd<C>, grid, dim3(kThreads), 0, stream, d));
    GOM_LAUNCH_CHECK();
    return GOM_OK;
}

int dispatch_tap(const GomLpipsTapArgs *p, bool bwd, cudaStream_t stream) {
    TapDev d;
    d.B = p->n_frames; d.h = p->height; d.w = p->width; d.pool = p->pool;
    d.feats = p->feats; d.lin = p->lin; d.layer_sums = p->layer_sums; d.pooled = p->pooled;
    d.dval = p->dL_dval; d.d_pooled = p->dL_dpooled; d.d_pre = p->dL_dpre;
    switch (p->channels) {
        case 32: return launch_tap<32>(d, bwd, stream);
        case 64: return launch_tap<64>(d, bwd, stream);
        case 128: return launch_tap<128>(d, bwd, stream);
        case 256: return launch_tap<256>(d, bwd, stream);
        case 512: return launch_tap<512>(d, bwd, stream);
        default:
            gom_set_error("gom_lpips_tap: channels must be 32, 64, 128, 256 or 512 (got %d)", p->channels);
            return GOM_ERR_UNSUPPORTED;
    }
}

}  // namespace

extern "C" int gom_lpips_input_forward(const GomLpipsInputArgs *p, gom_stream_t stream_) {
    GOM_REQUIRE(p != nullptr, "args");
    GOM_REQUIRE(p->n_frames > 0 && p->height > 0 && p->width > 0, "sizes");
    GOM_REQUIRE(p->pred && p->gt && p->out, "null pointer");
    cudaStream_t stream = (cudaStream_t)stream_;
    const long long n_half = 3LL * p->n_frames * p->height * p->width;
    gom_prof_begin(GOM_PROF_LPIPS_INPUT, stream);
    const bool vec = n_half % 12 == 0 && ((uintptr_t)p->pred % 16) == 0 && ((uintptr_t)p->gt % 16) == 0 && ((uintptr_t)p->out % 16) == 0;
    if (vec) {
        GOM_CUDA(gom_launch_pdl(k_lpips_input_fwd<true>, dim3(grid_for(2 * n_half / 12, kThreads)), dim3(kThreads), 0, stream,
                                p->pred, p->gt, p->out, n_half, p->from_unit_range ? 2.f : 1.f, p->from_unit_range ? -1.f : 0.f));
    } else {
        GOM_CUDA(gom_launch_pdl(k_lpips_input_fwd<false>, dim3(grid_for(2 * n_half, kThreads * 4)), dim3(kThreads), 0, stream,
                                p->pred, p->gt, p->out, n_half, p->from_unit_range ? 2.f : 1.f, p->from_unit_range ? -1.f : 0.f));
    }
    GOM_LAUNCH_CHECK();
    gom_prof_end(GOM_PROF_LPIPS_INPUT, stream);
    return GOM_OK;
}

extern "C" int gom_lpips_input_backward(const GomLpipsInputArgs *p, gom_stream_t stream_) {
    GOM_REQUIRE(p != nullptr, "args");
    GOM_REQUIRE(p->n_frames > 0 && p->height > 0 && p->width > 0, "sizes");
    GOM_REQUIRE(p->dL_dout && p->dL_dpred, "null pointer");
    cudaStream_t stream = (cudaStream_t)stream_;
    const long long n_half = 3LL * p->n_frames * p->height * p->width;
    gom_prof_begin(GOM_PROF_LPIPS_INPUT, stream);
    if (n_half % 12 == 0 && ((uintptr_t)p->dL_dout % 16) == 0 && ((uintptr_t)p->dL_dpred % 16) == 0) {
        GOM_CUDA(gom_launch_pdl(k_lpips_input_bwd<true>, dim3(grid_for(n_half / 12, kThreads)), dim3(kThreads), 0, stream, p->dL_dout, p->dL_dpred,
                                n_half, p->from_unit_range ? 2.f : 1.f));
    } else {
        GOM_CUDA(gom_launch_pdl(k_lpips_input_bwd<false>, dim3(grid_for(n_half, kThreads * 4)), dim3(kThreads), 0, stream, p->dL_dout, p->dL_dpred,
                                n_half, p->from_unit_range ? 2.f : 1.f));
    }
    GOM_LAUNCH_CHECK();
    gom_prof_end(GOM_PROF_LPIPS_INPUT, stream);
    return GOM_OK;
}

extern "C" int gom_bias_relu(const GomBiasReluArgs *p, gom_stream_t stream_) {
    GOM_REQUIRE(p != nullptr, "args");
    GOM_REQUIRE(p->n_pixels > 0 && p->channels > 0 && p->channels % 4 == 0, "channels must be a positive multiple of 4");
    GOM_REQUIRE(p->x && p->bias, "null pointer");
    GOM_REQUIRE(((uintptr_t)p->x % 16) == 0 && ((uintptr_t)p->bias % 16) == 0, "x / bias must be 16-byte aligned");
    cudaStream_t stream = (cudaStream_t)stream_;
    const long long n4 = p->n_pixels * (p->channels / 4);
    gom_prof_begin(GOM_PROF_BIAS_RELU, stream);
    k_bias_relu<<<grid_for(n4, kThreads * 4), kThreads, 0, stream>>>(reinterpret_cast<float4 *>(p->x),
                                                                    reinterpret_cast<const float4 *>(p->bias), n4, p->channels / 4);
    GOM_LAUNCH_CHECK();
    gom_prof_end(GOM_PROF_BIAS_RELU, stream);
    return GOM_OK;
}

extern "C" int gom_relu_backward(const GomReluBwdArgs *p, gom_stream_t stream_) {
    GOM_REQUIRE(p != nullptr, "args");
    GOM_REQUIRE(p->n > 0 && p->n % 4 == 0, "n must be a positive multiple of 4");
    GOM_REQUIRE(p->act && p->grad, "null pointer");
    GOM_REQUIRE(((uintptr_t)p->act % 16) == 0 && ((uintptr_t)p->grad % 16) == 0, "act / grad must be 16-byte aligned");
    cudaStream_t stream = (cudaStream_t)stream_;
    gom_prof_begin(GOM_PROF_RELU_BWD, stream);
    k_relu_bwd<<<grid_for(p->n / 4, kThreads * 4), kThreads, 0, stream>>>(reinterpret_cast<const float4 *>(p->act),
                                                                         reinterpret_cast<float4 *>(p->grad), p->n / 4);
    GOM_LAUNCH_CHECK();
    gom_prof_end(GOM_PROF_RELU_BWD, stream);
    return GOM_OK;
}

static int tap_common_checks(const GomLpipsTapArgs *p) {
    GOM_REQUIRE(p != nullptr, "args");
    GOM_REQUIRE(p->n_frames > 0 && p->n_frames <= 65535 && p->height > 0 && p->width > 0, "sizes");
    GOM_REQUIRE(p->feats && p->lin, "null pointer");
    GOM_REQUIRE(((uintptr_t)p->feats % 16) == 0, "feats must be 16-byte aligned");
    return GOM_OK;
}

extern "C" int gom_lpips_tap_forward(const GomLpipsTapArgs *p, gom_stream_t stream_) {
    if (int rc = tap_common_checks(p)) return rc;
    GOM_REQUIRE(p->layer_sums, "layer_sums");
    GOM_REQUIRE(!p->pool || (p->pooled && ((uintptr_t)p->pooled % 16) == 0), "pooled (16-byte aligned) is required when pool = 1");
    cudaStream_t stream = (cudaStream_t)stream_;
    gom_prof_begin(GOM_PROF_LPIPS_TAP_FWD, stream);
    const int rc = dispatch_tap(p, false, stream);
    gom_prof_end(GOM_PROF_LPIPS_TAP_FWD, stream);
    return rc;
}

extern "C" int gom_lpips_tap_backward(const GomLpipsTapArgs *p, gom_stream_t stream_) {
    if (int rc = tap_common_checks(p)) return rc;
    GOM_REQUIRE(p->dL_dval && p->dL_dpre && ((uintptr_t)p->dL_dpre % 16) == 0, "dL_dval / dL_dpre");
    GOM_REQUIRE(!p->dL_dpooled || ((uintptr_t)p->dL_dpooled % 16) == 0, "dL_dpooled alignment");
    cudaStream_t stream = (cudaStream_t)stream_;
    gom_prof_begin(GOM_PROF_LPIPS_TAP_BWD, stream);
    const int rc = dispatch_tap(p, true, stream);
    gom_prof_end(GOM_PROF_LPIPS_TAP_BWD, stream);
    return rc;
}

extern "C" size_t gom_sizeof_lpips_input_args(void) { return sizeof(GomLpipsInputArgs); }
extern "C" size_t gom_sizeof_bias_relu_args(void) { return sizeof(GomBiasReluArgs); }
extern "C" size_t gom_sizeof_relu_bwd_args(void) { return sizeof(GomReluBwdArgs); }
extern "C" size_t gom_sizeof_lpips_tap_args(void) { return sizeof(GomLpipsTapArgs); }
