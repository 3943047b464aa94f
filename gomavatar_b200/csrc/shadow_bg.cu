// shadow_bg.cu — the background row of the shadow MLP's backward pass (gom_shadow_mlp_background_*).
// Every background pixel of the normal map carries the same input (normal = 0), so the pseudo-shading MLP of reference
// models/modules/shadow_module.py:107-117 is evaluated for them ONCE (forward: the constant `bg_value` of csrc/shadow_mlp.cu).
// Their gradients are that one row's gradient scaled by the sum of their upstream gradients:
//   prepare: the row forward and backward for a unit upstream gradient (one block; thread i owns hidden unit i), then per pixel
//            g_normals[p] = [normal_p == 0] * g_out[p] * d out / d normal |_0  (foreground rows: 0, written by the tcgen05
//            backward afterwards) and g_bg = sum of g_out over the background pixels;
//   apply:   every parameter gradient += g_bg * (row gradient), formed on the fly from the row's dz / activation vectors.
// Three small launches instead of the ~80 torch kernels (matmuls, outer products, where / sum / add) that did this before.
#include <math.h>

#include "gom_common.cuh"

namespace {

constexpr int kWidth = 128, kMaxDepth = 8, kMaxIn = 64;
// scratch layout (floats)
constexpr int kMaxPixelBlocks = 148 * 8;
constexpr int S_GBG = 0, S_GX0 = 1, S_DZOUT = 4, S_COUNTER = 5, S_H = 8, S_DZ = S_H + kMaxDepth * kWidth, S_PART = S_DZ + kMaxDepth * kWidth,
              S_TOTAL = S_PART + kMaxPixelBlocks;

__global__ void __launch_bounds__(kWidth) k_shadow_bg_row(GomShadowMlpArgs a) {
    __shared__ float s_in[kMaxIn], s_h[kWidth], s_dz[kWidth], s_red[kWidth];
    const int i = threadIdx.x, in_dim = 3 + 6 * a.multires;
    float *sc = a.bg_scratch;
    // posenc(0) = [0, 0, 0, (sin 0, sin 0, sin 0, cos 0, cos 0, cos 0) per octave]
    if (i < kMaxIn) s_in[i] = (i >= 3 && i < in_dim && ((i - 3) % 6) >= 3) ? 1.f : 0.f;
    __syncthreads();
    // z = W v + b with a warp per output row (coalesced reads of the row-major weights), lanes over the inputs
    __shared__ float s_z[kWidth];
    const int warp = i >> 5, lane = i & 31;
    auto matvec = [&](const float *W, const float *bias, const float *v, int n_in) {
        for (int r = warp; r < kWidth; r += kWidth / 32) {
            float p = 0.f;
            for (int j = lane; j < n_in; j += 32) p += W[(long long)r * n_in + j] * v[j];
            p = warp_sum(p);
            if (lane == 0) s_z[r] = p + bias[r];
        }
        __syncthreads();
    };
    matvec(a.W_in, a.b_in, s_in, in_dim);
    float h = fmaxf(s_z[i], 0.f);
    sc[S_H + i] = h;
    for (int l = 1; l < a.depth; l++) {
        s_h[i] = h;
        __syncthreads();
        matvec(a.W_hid + (long long)(l - 1) * kWidth * kWidth, a.b_hid + (l - 1) * kWidth, s_h, kWidth);
        h = fmaxf(s_z[i], 0.f);
        sc[S_H + l * kWidth + i] = h;
    }
    // output unit: y0 = sigmoid(w_out . h + b_out), dz_out = y0 (1 - y0) for a unit upstream gradient
    s_red[i] = a.W_out[i] * h;
    __syncthreads();
    for (int d = kWidth / 2; d > 0; d >>= 1) {
        if (i < d) s_red[i] += s_red[i + d];
        __syncthreads();
    }
    const float y0 = 1.f / (1.f + expf(-(s_red[0] + a.b_out[0])));
    const float dz_out = y0 * (1.f - y0);
    __syncthreads();
    float dz = h > 0.f ? a.W_out[i] * dz_out : 0.f;                 // last hidden layer
    sc[S_DZ + (a.depth - 1) * kWidth + i] = dz;
    for (int l = a.depth - 1; l >= 1; l--) {                        // dz_{l-1} = (W_l^T dz_l) * [h_{l-1} > 0]
        s_dz[i] = dz;
        __syncthreads();
        const float *W = a.W_hid + (long long)(l - 1) * kWidth * kWidth;
        float acc = 0.f;
        for (int r = 0; r < kWidth; r++) acc += W[(long long)r * kWidth + i] * s_dz[r];
        __syncthreads();
        dz = sc[S_H + (l - 1) * kWidth + i] > 0.f ? acc : 0.f;
        sc[S_DZ + (l - 1) * kWidth + i] = dz;
    }
    // gradient w.r.t. the encoding, then through posenc at 0: d/dx = g_enc[d] + sum_k 2^k g_enc[3 + 6 k + d]  (cos 0 = 1, sin 0 = 0)
    s_dz[i] = dz;
    __syncthreads();
    if (i < in_dim) {
        float g = 0.f;
        for (int r = 0; r < kWidth; r++) g += a.W_in[r * in_dim + i] * s_dz[r];
        s_in[i] = g;
    }
    __syncthreads();
    if (i < 3) {
        float g = s_in[i], f = 1.f;
        for (int k = 0; k < a.multires; k++, f *= 2.f) g += f * s_in[3 + 6 * k + i];
        sc[S_GX0 + i] = g;
    }
    if (i == 0) { sc[S_GBG] = 0.f; sc[S_DZOUT] = dz_out; reinterpret_cast<unsigned int *>(sc)[S_COUNTER] = 0u; }
}

__global__ void __launch_bounds__(256) k_shadow_bg_pixels(GomShadowMlpArgs a) {
    __shared__ float s_part[8];
    const float *sc = a.bg_scratch;
    const float gx = sc[S_GX0], gy = sc[S_GX0 + 1], gz = sc[S_GX0 + 2];
    float sum = 0.f;
    for (long long p = (long long)blockIdx.x * 256 + threadIdx.x; p < a.n_pixels; p += (long long)gridDim.x * 256) {
        const float nx = a.normals[3 * p], ny = a.normals[3 * p + 1], nz = a.normals[3 * p + 2];
        const bool bg = nx == 0.f && ny == 0.f && nz == 0.f;
        const float g = bg ? a.g_out[p] : 0.f;
        sum += g;
        a.g_normals[3 * p] = g * gx; a.g_normals[3 * p + 1] = g * gy; a.g_normals[3 * p + 2] = g * gz;
    }
    sum = warp_sum(sum);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = sum;
    __syncthreads();
    // deterministic: per-block partial sums, added up in block order by whichever block finishes last
    __shared__ bool s_last;
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < 8; w++) s += s_part[w];
        a.bg_scratch[S_PART + blockIdx.x] = s;
        __threadfence();
        s_last = atomicAdd(reinterpret_cast<unsigned int *>(a.bg_scratch) + S_COUNTER, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        __threadfence();
        float s = 0.f;
        for (unsigned int b = 0; b < gridDim.x; b++) s += __ldcg(a.bg_scratch + S_PART + b);
        a.bg_scratch[S_GBG] = s;
    }
}

// every parameter gradient += g_bg * (unit-gradient row gradient)
__global__ void __launch_bounds__(256) k_shadow_bg_apply(GomShadowMlpArgs a) {
    const float *sc = a.bg_scratch;
    const float g_bg = sc[S_GBG];
    if (g_bg == 0.f) return;
    const int in_dim = 3 + 6 * a.multires;
    const long long n_in = (long long)kWidth * in_dim, n_hid = (long long)(a.depth - 1) * kWidth * kWidth;
    const long long total = n_in + n_hid + (long long)a.depth * kWidth + kWidth + 1;
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
        if (e < n_in) {                                                   // dW_in = dz_0 (x) posenc(0)
            const int i = (int)(e / in_dim), j = (int)(e % in_dim);
            const float enc = (j >= 3 && ((j - 3) % 6) >= 3) ? 1.f : 0.f;
            if (enc != 0.f) a.g_W_in[e] += g_bg * sc[S_DZ + i];
        } else if (e < n_in + n_hid) {                                     // dW_l = dz_l (x) h_{l-1}
            const long long q = e - n_in;
            const int l = (int)(q / (kWidth * kWidth)) + 1, r = (int)(q % (kWidth * kWidth));
            a.g_W_hid[q] += g_bg * sc[S_DZ + l * kWidth + r / kWidth] * sc[S_H + (l - 1) * kWidth + r % kWidth];
        } else if (e < n_in + n_hid + (long long)a.depth * kWidth) {      // biases: dz_l
            const int q = (int)(e - n_in - n_hid), l = q / kWidth, i = q % kWidth;
            if (l == 0) a.g_b_in[i] += g_bg * sc[S_DZ + i];
            else a.g_b_hid[(l - 1) * kWidth + i] += g_bg * sc[S_DZ + l * kWidth + i];
        } else if (e < total - 1) {                                        // dw_out = dz_out h_last
            const int i = (int)(e - n_in - n_hid - (long long)a.depth * kWidth);
            a.g_w_out[i] += g_bg * sc[S_DZOUT] * sc[S_H + (a.depth - 1) * kWidth + i];
        } else {
            a.g_b_out[0] += g_bg * sc[S_DZOUT];
        }
    }
}

int check(const GomShadowMlpArgs *p) {
    GOM_REQUIRE(p != nullptr, "args");
    GOM_REQUIRE(p->n_pixels > 0, "n_pixels");
    GOM_REQUIRE(p->width == kWidth, "width must be 128");
    GOM_REQUIRE(p->depth >= 1 && p->depth <= kMaxDepth, "depth must be in [1, 8]");
    GOM_REQUIRE(p->multires >= 0 && p->multires <= 10, "multires must be in [0, 10]");
    GOM_REQUIRE(p->W_in && p->b_in && p->W_out && p->b_out && (p->depth == 1 || (p->W_hid && p->b_hid)), "null weights");
    GOM_REQUIRE(p->bg_scratch, "null bg_scratch");
    return GOM_OK;
}

}  // namespace

extern "C" size_t gom_shadow_mlp_bg_scratch_floats(void) { return (size_t)S_TOTAL; }

extern "C" int gom_shadow_mlp_background_prepare(const GomShadowMlpArgs *p, gom_stream_t stream_) {
    if (int rc = check(p)) return rc;
    GOM_REQUIRE(p->normals && p->g_out && p->g_normals, "null pixel buffers");
    cudaStream_t stream = (cudaStream_t)stream_;
    k_shadow_bg_row<<<1, kWidth, 0, stream>>>(*p);
    GOM_LAUNCH_CHECK();
    int blocks = gom_div_up(p->n_pixels, 256 * 8);
    if (blocks > kMaxPixelBlocks) blocks = kMaxPixelBlocks;
    k_shadow_bg_pixels<<<blocks, 256, 0, stream>>>(*p);
    GOM_LAUNCH_CHECK();
    return GOM_OK;
}

extern "C" int gom_shadow_mlp_background_apply(const GomShadowMlpArgs *p, gom_stream_t stream_) {
    if (int rc = check(p)) return rc;
    GOM_REQUIRE(p->g_W_in && p->g_b_in && p->g_w_out && p->g_b_out && (p->depth == 1 || (p->g_W_hid && p->g_b_hid)), "null gradient buffers");
    k_shadow_bg_apply<<<64, 256, 0, (cudaStream_t)stream_>>>(*p);
    GOM_LAUNCH_CHECK();
    return GOM_OK;
}
