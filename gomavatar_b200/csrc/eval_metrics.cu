// eval_metrics.cu — the reference's evaluation metrics on the GPU, bit-faithful to its CPU definitions (sm_100a).
//
// Replaces reference eval.py:101-108 (Evaluator.psnr_metric / ssim_metric) together with the 8-bit quantisation
// that precedes them (utils/image_util.py:21-22 `to_8b_image`, applied at eval.py:355-361):
//   img8 = uint8(255 * clip(img, 0, 1))   (float32 product, truncation)
//   mse  = mean((p - g)^2), p = img8 / 255 in float64;   psnr = -10 ln(mse) / ln(10)
//   ssim = skimage 0.18 `structural_similarity(p, g, multichannel=True)` defaults: 7x7 uniform window, sample
//          covariance (49/48), K1 = .01, K2 = .03, data_range = 2 for float input, 3-pixel border cropped, mean over
//          pixels and channels.
// Because p and g are integers / 255, every window sum (x, y, x^2, y^2, xy over 49 pixels) is computed EXACTLY in
// int32 and only the final SSIM expression is evaluated in float64 — closer to real arithmetic than skimage's own
// float64 running sums, and within 1e-12 of them.
#include "gom_common.cuh"

namespace {

constexpr int kWin = 7, kPad = 3;
constexpr int kTx = 32, kTy = 8;                       // output pixels per block
constexpr int kSx = kTx + 2 * kPad, kSy = kTy + 2 * kPad;

struct EvalDev {
    int B, H, W, quantize;
    const float *pred, *gt;
    double *ssim_sum; unsigned long long *sq_err_sum; unsigned char *pred_8b;
};

__device__ __forceinline__ int to_8b(float v, int quantize) {
    if (quantize) return (int)(unsigned char)(255.f * fminf(fmaxf(v, 0.f), 1.f));      // utils/image_util.py:21-22
    return (int)rintf(255.f * fminf(fmaxf(v, 0.f), 1.f));                              // input already is k / 255
}

__global__ void __launch_bounds__(kTx * kTy) k_eval_metrics(EvalDev a) {
    __shared__ unsigned char sp[3][kSy][kSx + 2], sg[3][kSy][kSx + 2];
    __shared__ double red_s[kTx * kTy / 32];
    __shared__ unsigned long long red_e[kTx * kTy / 32];
    const int b = blockIdx.z;
    const int ox = blockIdx.x * kTx, oy = blockIdx.y * kTy;           // tile origin in image coordinates
    const int tid = threadIdx.y * kTx + threadIdx.x;
    const float *P = a.pred + (long long)b * a.H * a.W * 3, *G = a.gt + (long long)b * a.H * a.W * 3;
    unsigned long long sq = 0;
    // stage the tile + 3-pixel halo, quantised; the squared error is counted on the tile's own pixels only
    for (int i = tid; i < kSy * kSx; i += kTx * kTy) {
        const int sy = i / kSx, sx = i - sy * kSx;
        const int y = oy + sy - kPad, x = ox + sx - kPad;
        int p[3] = {0, 0, 0}, g[3] = {0, 0, 0};
        if (y >= 0 && y < a.H && x >= 0 && x < a.W) {
            const long long o = ((long long)y * a.W + x) * 3;
#pragma unroll
            for (int c = 0; c < 3; c++) { p[c] = to_8b(P[o + c], a.quantize); g[c] = to_8b(G[o + c], a.quantize); }
            if (sy >= kPad && sy < kPad + kTy && sx >= kPad && sx < kPad + kTx) {
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    const int d = p[c] - g[c];
                    sq += (unsigned long long)(d * d);
                    if (a.pred_8b) a.pred_8b[((long long)b * a.H * a.W) * 3 + o + c] = (unsigned char)p[c];
                }
            }
        }
#pragma unroll
        for (int c = 0; c < 3; c++) { sp[c][sy][sx] = (unsigned char)p[c]; sg[c][sy][sx] = (unsigned char)g[c]; }
    }
    __syncthreads();
    const int x = ox + threadIdx.x, y = oy + threadIdx.y;
    double s_sum = 0.0;
    if (x >= kPad && x < a.W - kPad && y >= kPad && y < a.H - kPad) {          // windows that fit (the crop keeps these)
        const double NP = kWin * kWin, cov_norm = NP / (NP - 1.0);
        const double C1 = (0.01 * 2.0) * (0.01 * 2.0), C2 = (0.03 * 2.0) * (0.03 * 2.0);
#pragma unroll
        for (int c = 0; c < 3; c++) {
            int sx_ = 0, sy_ = 0, sxx = 0, syy = 0, sxy = 0;
            for (int dy = 0; dy < kWin; dy++)
#pragma unroll
                for (int dx = 0; dx < kWin; dx++) {
                    const int p = sp[c][threadIdx.y + dy][threadIdx.x + dx], g = sg[c][threadIdx.y + dy][threadIdx.x + dx];
                    sx_ += p; sy_ += g; sxx += p * p; syy += g * g; sxy += p * g;
                }
            const double ux = sx_ / (NP * 255.0), uy = sy_ / (NP * 255.0);
            const double uxx = sxx / (NP * 65025.0), uyy = syy / (NP * 65025.0), uxy = sxy / (NP * 65025.0);
            const double vx = cov_norm * (uxx - ux * ux), vy = cov_norm * (uyy - uy * uy), vxy = cov_norm * (uxy - ux * uy);
            s_sum += ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux * ux + uy * uy + C1) * (vx + vy + C2));
        }
    }
    // block reduction -> one fp64 / one u64 atomic per block
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        s_sum += __shfl_xor_sync(0xffffffffu, s_sum, d);
        sq += __shfl_xor_sync(0xffffffffu, sq, d);
    }
    if ((tid & 31) == 0) { red_s[tid >> 5] = s_sum; red_e[tid >> 5] = sq; }
    __syncthreads();
    if (tid == 0) {
        double s = 0.0; unsigned long long e = 0;
        for (int k = 0; k < kTx * kTy / 32; k++) { s += red_s[k]; e += red_e[k]; }
        atomicAdd(a.ssim_sum + b, s);
        atomicAdd(a.sq_err_sum + b, e);
    }
}

}  // namespace

extern "C" int gom_eval_metrics(const GomEvalMetricsArgs *p, gom_stream_t stream_) {
    GOM_REQUIRE(p != nullptr, "args");
    GOM_REQUIRE(p->n_frames > 0 && p->n_frames <= 65535 && p->height >= kWin && p->width >= kWin, "sizes (image must be at least 7x7)");
    GOM_REQUIRE(p->pred && p->gt && p->ssim_sum && p->sq_err_sum, "null pointer");
    cudaStream_t stream = (cudaStream_t)stream_;
    EvalDev a;
    a.B = p->n_frames; a.H = p->height; a.W = p->width; a.quantize = p->quantize;
    a.pred = p->pred; a.gt = p->gt; a.ssim_sum = p->ssim_sum;
    a.sq_err_sum = reinterpret_cast<unsigned long long *>(p->sq_err_sum); a.pred_8b = p->pred_8b;
    GOM_CUDA(cudaMemsetAsync(a.ssim_sum, 0, sizeof(double) * a.B, stream));
    GOM_CUDA(cudaMemsetAsync(a.sq_err_sum, 0, sizeof(unsigned long long) * a.B, stream));
    dim3 grid(gom_div_up(a.W, kTx), gom_div_up(a.H, kTy), a.B), block(kTx, kTy);
    gom_prof_begin(GOM_PROF_EVAL_METRICS, stream);
    k_eval_metrics<<<grid, block, 0, stream>>>(a);
    GOM_LAUNCH_CHECK();
    gom_prof_end(GOM_PROF_EVAL_METRICS, stream);
    return GOM_OK;
}

extern "C" size_t gom_sizeof_eval_metrics_args(void) { return sizeof(GomEvalMetricsArgs); }
