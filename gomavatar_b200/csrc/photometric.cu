// photometric.cu — background compositing + L1 photometric losses, forward and backward, fused (sm_100a).
//
// Replaces reference train.py:53-55 (`unpack`: rgb*mask + bg*(1-mask)) and train.py:101-111 (L1 on rgb and on the
// mask) — 10 elementwise/reduction launches forward and as many in autograd — with one pass each way.  The unpacked
// image is also an output because LPIPS (train.py:113-121) consumes it; its gradient comes back as g_unpacked.
#include <algorithm>

#include "gom_common.cuh"

namespace {
constexpr int kThreads = 256;

struct PhotoDev {
    long long n_pix_per_frame; int B;
    const float *rgb; long long rgb_ps;      // pixel stride in floats (3 or 4)
    const float *mask; long long mask_ps;    // pixel stride in floats (1 or 4)
    const float *bg;                         // [B,3] (nullable: no compositing, unpacked = rgb)
    const float *gt_rgb, *gt_mask;           // [B,H,W,3], [B,H,W]
    float *unpacked;                         // [B,H,W,3]
    float *loss_sums;                        // [2]: sum|u-gt|, sum|m-gt_m|
    const float *g_unpacked;                 // [B,H,W,3] nullable
    const float *g_loss;                     // [2] device: dL/d(loss_rgb), dL/d(loss_mask) (already mean-scaled by host? no: raw)
    float inv_n_rgb, inv_n_mask;
    float *d_rgb; long long d_rgb_ps;        // gradient wrt rgb, same pixel stride convention
    float *d_mask; long long d_mask_ps;
    int rgba4, d_rgba4;                      // rgb | mask (resp. their gradients) are the channels of ONE 16-byte aligned [.,4] tensor
};

__device__ __forceinline__ int frame_of(long long i, long long n_pix_per_frame, long long total) {
    return total < (1ll << 31) ? (int)((uint32_t)i / (uint32_t)n_pix_per_frame) : (int)(i / n_pix_per_frame);
}

__device__ __forceinline__ float sgn(float x) { return (x > 0.f) - (x < 0.f); }

__global__ void __launch_bounds__(kThreads) k_photo_fwd(PhotoDev a) {
    const long long total = a.n_pix_per_frame * a.B;
    float s_rgb = 0.f, s_mask = 0.f;
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
        const int b = frame_of(i, a.n_pix_per_frame, total);
        float m, u[3];
        if (a.rgba4) {
            const float4 f = __ldg(reinterpret_cast<const float4 *>(a.rgb) + i);
            u[0] = f.x; u[1] = f.y; u[2] = f.z; m = f.w;
        } else {
            m = a.mask[i * a.mask_ps];
            const float *c = a.rgb + i * a.rgb_ps;
            u[0] = c[0]; u[1] = c[1]; u[2] = c[2];
        }
        if (a.bg) {
#pragma unroll
            for (int k = 0; k < 3; k++) u[k] = u[k] * m + a.bg[3 * b + k] * (1.f - m);
        }
#pragma unroll
        for (int k = 0; k < 3; k++) {
            a.unpacked[3 * i + k] = u[k];
            if (a.gt_rgb) s_rgb += fabsf(u[k] - a.gt_rgb[3 * i + k]);
        }
        if (a.gt_mask) s_mask += fabsf(m - a.gt_mask[i]);
    }
    if (!a.loss_sums) return;
    __shared__ float sh[2][kThreads / 32];
    s_rgb = warp_sum(s_rgb); s_mask = warp_sum(s_mask);
    if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s_rgb; sh[1][threadIdx.x >> 5] = s_mask; }
    __syncthreads();
    if (threadIdx.x < 32) {
        float x = threadIdx.x < kThreads / 32 ? sh[0][threadIdx.x] : 0.f, y = threadIdx.x < kThreads / 32 ? sh[1][threadIdx.x] : 0.f;
        x = warp_sum(x); y = warp_sum(y);
        if (threadIdx.x == 0) { atomicAdd(a.loss_sums, x); atomicAdd(a.loss_sums + 1, y); }
    }
}

__global__ void __launch_bounds__(kThreads) k_photo_bwd(PhotoDev a) {
    const long long total = a.n_pix_per_frame * a.B;
    const float w_rgb = a.g_loss ? a.g_loss[0] * a.inv_n_rgb : 0.f, w_mask = a.g_loss ? a.g_loss[1] * a.inv_n_mask : 0.f;
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
        const int b = frame_of(i, a.n_pix_per_frame, total);
        float m, c[3], d[3];
        if (a.rgba4) {
            const float4 f = __ldg(reinterpret_cast<const float4 *>(a.rgb) + i);
            c[0] = f.x; c[1] = f.y; c[2] = f.z; m = f.w;
        } else {
            m = a.mask[i * a.mask_ps];
            const float *cp = a.rgb + i * a.rgb_ps;
            c[0] = cp[0]; c[1] = cp[1]; c[2] = cp[2];
        }
        float dm = (a.gt_mask && a.g_loss) ? w_mask * sgn(m - a.gt_mask[i]) : 0.f;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float bgk = a.bg ? a.bg[3 * b + k] : 0.f;
            const float u = a.bg ? c[k] * m + bgk * (1.f - m) : c[k];
            float gu = a.g_unpacked ? a.g_unpacked[3 * i + k] : 0.f;
            if (a.gt_rgb && a.g_loss) gu += w_rgb * sgn(u - a.gt_rgb[3 * i + k]);
            if (a.bg) {
                d[k] = gu * m;
                dm += gu * (c[k] - bgk);
            } else {
                d[k] = gu;
            }
        }
        if (a.d_rgba4) {
            reinterpret_cast<float4 *>(a.d_rgb)[i] = make_float4(d[0], d[1], d[2], dm);
        } else {
#pragma unroll
            for (int k = 0; k < 3; k++) a.d_rgb[i * a.d_rgb_ps + k] = d[k];
            a.d_mask[i * a.d_mask_ps] = dm;
        }
    }
}
}  // namespace

static int photo_fill(const GomPhotoArgs *p, PhotoDev &a) {
    GOM_REQUIRE(p != nullptr, "args");
    GOM_REQUIRE(p->n_frames > 0 && p->height > 0 && p->width > 0, "sizes");
    GOM_REQUIRE(p->rgb && p->mask, "null input");
    GOM_REQUIRE(p->rgb_pixel_stride >= 3 && p->mask_pixel_stride >= 1, "pixel strides");
    a.B = p->n_frames; a.n_pix_per_frame = (long long)p->height * p->width;
    a.rgb = p->rgb; a.rgb_ps = p->rgb_pixel_stride; a.mask = p->mask; a.mask_ps = p->mask_pixel_stride;
    a.bg = p->bgcolor; a.gt_rgb = p->gt_rgb; a.gt_mask = p->gt_mask; a.unpacked = p->unpacked; a.loss_sums = p->loss_sums;
    a.g_unpacked = p->dL_dunpacked; a.g_loss = p->dL_dlosses;
    a.inv_n_rgb = 1.0f / (3.0f * (float)a.n_pix_per_frame * a.B); a.inv_n_mask = 1.0f / ((float)a.n_pix_per_frame * a.B);
    a.d_rgb = p->dL_drgb; a.d_rgb_ps = p->dL_drgb_pixel_stride; a.d_mask = p->dL_dmask; a.d_mask_ps = p->dL_dmask_pixel_stride;
    a.rgba4 = a.rgb_ps == 4 && a.mask_ps == 4 && a.mask == a.rgb + 3 && ((uintptr_t)a.rgb % 16) == 0;
    a.d_rgba4 = a.d_rgb && a.d_rgb_ps == 4 && a.d_mask_ps == 4 && a.d_mask == a.d_rgb + 3 && ((uintptr_t)a.d_rgb % 16) == 0;
    return GOM_OK;
}

extern "C" int gom_photometric_forward(const GomPhotoArgs *p, gom_stream_t stream_) {
    PhotoDev a{};
    int rc = photo_fill(p, a);
    if (rc) return rc;
    GOM_REQUIRE(p->unpacked, "null output");
    cudaStream_t stream = (cudaStream_t)stream_;
    if (a.loss_sums) GOM_CUDA(cudaMemsetAsync(a.loss_sums, 0, 2 * sizeof(float), stream));
    const long long total = a.n_pix_per_frame * a.B;
    const int grid = (int)std::min<long long>((total + kThreads - 1) / kThreads, 148 * 8);
    gom_prof_begin(GOM_PROF_PHOTO_FWD, stream);
    k_photo_fwd<<<grid, kThreads, 0, stream>>>(a);
    GOM_LAUNCH_CHECK();
    gom_prof_end(GOM_PROF_PHOTO_FWD, stream);
    return GOM_OK;
}

extern "C" int gom_photometric_backward(const GomPhotoArgs *p, gom_stream_t stream_) {
    PhotoDev a{};
    int rc = photo_fill(p, a);
    if (rc) return rc;
    GOM_REQUIRE(p->dL_drgb && p->dL_dmask, "null output");
    GOM_REQUIRE(p->dL_drgb_pixel_stride >= 3 && p->dL_dmask_pixel_stride >= 1, "gradient pixel strides");
    const long long total = a.n_pix_per_frame * a.B;
    const int grid = (int)std::min<long long>((total + kThreads - 1) / kThreads, 148 * 8);
    gom_prof_begin(GOM_PROF_PHOTO_BWD, (cudaStream_t)stream_);
    k_photo_bwd<<<grid, kThreads, 0, (cudaStream_t)stream_>>>(a);
    GOM_LAUNCH_CHECK();
    gom_prof_end(GOM_PROF_PHOTO_BWD, (cudaStream_t)stream_);
    return GOM_OK;
}

// ------------------------------------------------------------------------------------------------ pseudo-shading
namespace {
__global__ void __launch_bounds__(kThreads) k_shade_fwd(GomShadeArgs a) {
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < a.n_pixels; i += (long long)gridDim.x * kThreads) {
        const float4 f = __ldg(reinterpret_cast<const float4 *>(a.rgba) + i);
        const float s = __ldg(a.shading + i);
        a.rgbs[3 * i] = f.x * s; a.rgbs[3 * i + 1] = f.y * s; a.rgbs[3 * i + 2] = f.z * s;
        a.masks[i] = f.w;
    }
}
__global__ void __launch_bounds__(kThreads) k_shade_bwd(GomShadeArgs a) {
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < a.n_pixels; i += (long long)gridDim.x * kThreads) {
        const float4 f = __ldg(reinterpret_cast<const float4 *>(a.rgba) + i);
        const float s = __ldg(a.shading + i);
        float g0 = 0.f, g1 = 0.f, g2 = 0.f;
        if (a.dL_drgbs) { g0 = a.dL_drgbs[3 * i]; g1 = a.dL_drgbs[3 * i + 1]; g2 = a.dL_drgbs[3 * i + 2]; }
        const float gm = a.dL_dmasks ? a.dL_dmasks[i] : 0.f;
        reinterpret_cast<float4 *>(a.dL_drgba)[i] = make_float4(g0 * s, g1 * s, g2 * s, gm);
        a.dL_dshading[i] = g0 * f.x + g1 * f.y + g2 * f.z;
    }
}
}  // namespace

extern "C" int gom_shade_forward(const GomShadeArgs *p, gom_stream_t stream_) {
    GOM_REQUIRE(p != nullptr, "args");
    GOM_REQUIRE(p->n_pixels > 0, "sizes");
    GOM_REQUIRE(p->rgba && p->shading && p->rgbs && p->masks, "null pointer");
    GOM_REQUIRE(((uintptr_t)p->rgba % 16) == 0, "rgba must be 16-byte aligned");
    const int grid = (int)std::min<long long>((p->n_pixels + kThreads - 1) / kThreads, 148 * 8);
    k_shade_fwd<<<grid, kThreads, 0, (cudaStream_t)stream_>>>(*p);
    GOM_LAUNCH_CHECK();
    return GOM_OK;
}

extern "C" int gom_shade_backward(const GomShadeArgs *p, gom_stream_t stream_) {
    GOM_REQUIRE(p != nullptr, "args");
    GOM_REQUIRE(p->n_pixels > 0, "sizes");
    GOM_REQUIRE(p->rgba && p->shading && p->dL_drgba && p->dL_dshading, "null pointer");
    GOM_REQUIRE(((uintptr_t)p->rgba % 16) == 0 && ((uintptr_t)p->dL_drgba % 16) == 0, "rgba / dL_drgba must be 16-byte aligned");
    const int grid = (int)std::min<long long>((p->n_pixels + kThreads - 1) / kThreads, 148 * 8);
    k_shade_bwd<<<grid, kThreads, 0, (cudaStream_t)stream_>>>(*p);
    GOM_LAUNCH_CHECK();
    return GOM_OK;
}

extern "C" size_t gom_sizeof_shade_args(void) { return sizeof(GomShadeArgs); }
