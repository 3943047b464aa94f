// conv_first.cu — the first VGG16 convolution of LPIPS (3 -> 64 channels, 3x3, pad 1) with bias + ReLU, and its
// input gradient, hand-written for sm_100a.
//
// Replaces, for reference utils/lpips/pretrained_networks.py:96-134 (`vgg16().features[0:2]`, slice1's conv1_1 +
// ReLU) what cuDNN runs on B200 as a pre-Blackwell `sm80_xmma_fprop_implicit_gemm_indexed_wo_smem_tf32` kernel plus
// two layout-conversion helpers (0.5 ms per 16 images at 512x512 — as long as conv1_2, which has 21x the FLOPs; see
// profiles/r1s_launches_step.md) and a 0.4 ms TF32 dgrad.  With K = 27 the layer is not tensor-core work: it is bound
// by writing 64 channels per pixel (forward) / reading them (backward).  Here it is exact fp32 on the CUDA cores:
//   forward : lane = output-channel pair, weights live in registers, a 3x3x3 input window slides along the row in
//             registers (9 shared-memory broadcasts per pixel), 27 packed FFMA2 (Blackwell `fma.rn.f32x2`) per pixel,
//             one coalesced 256-byte store per pixel and warp;
//   backward: thread = 8 adjacent input pixels, gradient tile staged channel-major in shared memory in chunks of 8
//             output channels (16-byte loads), weights broadcast from shared memory.
#include "gom_common.cuh"

namespace {

constexpr int kCout = 64, kCin = 3;

__device__ __forceinline__ void ffma2(float2 &acc, const float2 w, const float2 x) {
    unsigned long long a = *reinterpret_cast<unsigned long long *>(&acc);
    const unsigned long long b = *reinterpret_cast<const unsigned long long *>(&w);
    const unsigned long long c = *reinterpret_cast<const unsigned long long *>(&x);
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a) : "l"(b), "l"(c));
    acc = *reinterpret_cast<float2 *>(&a);
}

// ------------------------------------------------------------------------------------------------------- forward
constexpr int kFwdRows = 8, kFwdCols = 64;                     // output tile per 256-thread block (one row per warp)
constexpr int kFwdSW = kFwdCols + 2;

struct ConvFwdDev { int N, H, W; const float *x, *w, *bias; float *out; };

// column c of the 3-row window: [row][ci]
__device__ __forceinline__ void load_col(const float (*tile)[kFwdSW][kCin], int row0, int c, float (&col)[9]) {
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int ci = 0; ci < kCin; ci++) col[r * 3 + ci] = tile[row0 + r][c][ci];
}

__device__ __forceinline__ float2 dup(float v) { return make_float2(v, v); }

// two adjacent output pixels from four input columns: four independent FFMA2 chains keep the FMA pipe fed
__device__ __forceinline__ void pixel2(const float (&c0)[9], const float (&c1)[9], const float (&c2)[9], const float (&c3)[9],
                                       const float2 (&w)[27], float2 bias, float2 &o0, float2 &o1) {
    float2 a0 = bias, a1 = make_float2(0.f, 0.f), b0 = bias, b1 = make_float2(0.f, 0.f);   // w index: (ky*3+kx)*3+ci
#pragma unroll
    for (int ky = 0; ky < 3; ky++)
#pragma unroll
        for (int ci = 0; ci < kCin; ci++) {
            const float2 w0 = w[(ky * 3 + 0) * 3 + ci], w1 = w[(ky * 3 + 1) * 3 + ci], w2 = w[(ky * 3 + 2) * 3 + ci];
            const float2 v0 = dup(c0[ky * 3 + ci]), v1 = dup(c1[ky * 3 + ci]), v2 = dup(c2[ky * 3 + ci]), v3 = dup(c3[ky * 3 + ci]);
            ffma2(a0, w0, v0); ffma2(b0, w0, v1);
            ffma2(a1, w1, v1); ffma2(b1, w1, v2);
            ffma2(a0, w2, v2); ffma2(b0, w2, v3);
        }
    o0 = make_float2(fmaxf(a0.x + a1.x, 0.f), fmaxf(a0.y + a1.y, 0.f));
    o1 = make_float2(fmaxf(b0.x + b1.x, 0.f), fmaxf(b0.y + b1.y, 0.f));
}

constexpr int kTileElems = (kFwdRows + 2) * kFwdSW * kCin;       // 1980 input values per tile
constexpr int kTilePerThread = (kTileElems + 255) / 256;         // 8

// A block owns an 8-row band of one image and walks it in 64-column chunks: the weights are loaded into registers once,
// the next chunk's input tile is fetched into registers while the current one is convolved (double-buffered smem).
__global__ void __launch_bounds__(256, 2) k_conv_first_fwd(ConvFwdDev a) {
    __shared__ float tile[2][kFwdRows + 2][kFwdSW][kCin];        // zero-padded input tile
    const int n = blockIdx.y, y0 = blockIdx.x * kFwdRows;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const float *X = a.x + (long long)n * a.H * a.W * kCin;
    float stage[kTilePerThread];
    auto fetch = [&](int x0) {
#pragma unroll
        for (int k = 0; k < kTilePerThread; k++) {
            const int i = threadIdx.x + k * 256;
            const int ci = i % kCin, c = (i / kCin) % kFwdSW, r = i / (kCin * kFwdSW);
            const int y = y0 + r - 1, x = x0 + c - 1;
            stage[k] = (i < kTileElems && y >= 0 && y < a.H && x >= 0 && x < a.W) ? X[((long long)y * a.W + x) * kCin + ci] : 0.f;
        }
    };
    auto commit = [&](int buf) {
        float *t = &tile[buf][0][0][0];
#pragma unroll
        for (int k = 0; k < kTilePerThread; k++) {
            const int i = threadIdx.x + k * 256;
            if (i < kTileElems) t[i] = stage[k];
        }
    };
    fetch(0);
    // this lane's two output channels: weights [Cout][Cin][3][3] -> w[(ky*3+kx)*3+ci] = (w[2l], w[2l+1])
    float2 w[27];
#pragma unroll
    for (int ky = 0; ky < 3; ky++)
#pragma unroll
        for (int kx = 0; kx < 3; kx++)
#pragma unroll
            for (int ci = 0; ci < kCin; ci++) {
                const int o = (ci * 3 + ky) * 3 + kx;
                w[(ky * 3 + kx) * 3 + ci] = make_float2(__ldg(a.w + (2 * lane) * 27 + o), __ldg(a.w + (2 * lane + 1) * 27 + o));
            }
    const float2 bias = make_float2(__ldg(a.bias + 2 * lane), __ldg(a.bias + 2 * lane + 1));
    commit(0);
    __syncthreads();
    const int y = y0 + wid;
    int buf = 0;
    for (int x0 = 0; x0 < a.W; x0 += kFwdCols, buf ^= 1) {
        const bool more = x0 + kFwdCols < a.W;
        if (more) fetch(x0 + kFwdCols);                           // global loads in flight during the convolution below
        if (y < a.H) {
            const float (*tl)[kFwdSW][kCin] = tile[buf];
            float2 *O = reinterpret_cast<float2 *>(a.out + (((long long)n * a.H + y) * a.W) * kCout) + lane;
            const int ncols = min(kFwdCols, a.W - x0);
            float A[9], B[9], C[9], D[9];
            float2 o0, o1;
            load_col(tl, wid, 0, A);
            load_col(tl, wid, 1, B);
            int c = 0;
            for (; c + 4 <= ncols; c += 4) {                      // window rotates through (A,B | C,D) without moves
                load_col(tl, wid, c + 2, C);
                load_col(tl, wid, c + 3, D);
                pixel2(A, B, C, D, w, bias, o0, o1);
                O[(long long)(x0 + c) * (kCout / 2)] = o0;
                O[(long long)(x0 + c + 1) * (kCout / 2)] = o1;
                load_col(tl, wid, c + 4, A);
                load_col(tl, wid, c + 5, B);
                pixel2(C, D, A, B, w, bias, o0, o1);
                O[(long long)(x0 + c + 2) * (kCout / 2)] = o0;
                O[(long long)(x0 + c + 3) * (kCout / 2)] = o1;
            }
            for (; c < ncols; c += 2) {                           // ragged right edge (tile columns beyond are zero / unused)
                load_col(tl, wid, c + 2, C);
                load_col(tl, wid, min(c + 3, kFwdSW - 1), D);
                pixel2(A, B, C, D, w, bias, o0, o1);
                O[(long long)(x0 + c) * (kCout / 2)] = o0;
                if (c + 1 < ncols) O[(long long)(x0 + c + 1) * (kCout / 2)] = o1;
#pragma unroll
                for (int k = 0; k < 9; k++) { A[k] = C[k]; B[k] = D[k]; }
            }
        }
        if (more) commit(buf ^ 1);
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------------- backward
// dx[y,x,ci] = sum_{ky,kx,co} g[y+1-ky, x+1-kx, co] * w[co,ci,ky,kx]
constexpr int kBwdRows = 16, kBwdCols = 64, kBwdPx = 8, kCoChunk = 8;
constexpr int kBwdThreads = kBwdRows * kBwdCols / kBwdPx;      // 128
constexpr int kBwdSW = kBwdCols + 4;                           // 1 halo column each side, padded to a multiple of 4

struct ConvBwdDev { int N, H, W; const float *g, *w; float *dx; };

__global__ void __launch_bounds__(kBwdThreads) k_conv_first_bwd(ConvBwdDev a) {
    __shared__ __align__(16) float sg[kCoChunk][kBwdRows + 2][kBwdSW];   // gradient tile, channel-major; col 0 = x0-1
    __shared__ __align__(16) float sw[kCout][9][4];                      // w[co][ky*3+kx][ci] (+ pad)
    const int n = blockIdx.z, y0 = blockIdx.y * kBwdRows, x0 = blockIdx.x * kBwdCols;
    const float *G = a.g + (long long)n * a.H * a.W * kCout;
    for (int i = threadIdx.x; i < kCout * 9 * 4; i += kBwdThreads) {
        const int ci = i & 3, t = (i >> 2) % 9, co = i / 36;
        sw[co][t][ci] = ci < 3 ? a.w[(co * 3 + ci) * 9 + t] : 0.f;
    }
    const int ty = threadIdx.x / (kBwdCols / kBwdPx), tx = (threadIdx.x % (kBwdCols / kBwdPx)) * kBwdPx;
    float acc[kBwdPx][3];
#pragma unroll
    for (int p = 0; p < kBwdPx; p++) acc[p][0] = acc[p][1] = acc[p][2] = 0.f;
    constexpr int kTilePix = (kBwdRows + 2) * (kBwdCols + 2), kVecPerPix = kCoChunk / 4;
    for (int c0 = 0; c0 < kCout; c0 += kCoChunk) {
        __syncthreads();
        // stage g[y0-1 .. y0+16][x0-1 .. x0+64][c0 .. c0+7] with 16-byte loads, transposed to channel-major
        for (int i = threadIdx.x; i < kTilePix * kVecPerPix; i += kBwdThreads) {
            const int q = i % kVecPerPix, pixi = i / kVecPerPix;
            const int c = pixi % (kBwdCols + 2), r = pixi / (kBwdCols + 2);
            const int y = y0 + r - 1, x = x0 + c - 1;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (y >= 0 && y < a.H && x >= 0 && x < a.W)
                v = *reinterpret_cast<const float4 *>(G + ((long long)y * a.W + x) * kCout + c0 + 4 * q);
            sg[4 * q + 0][r][c] = v.x; sg[4 * q + 1][r][c] = v.y; sg[4 * q + 2][r][c] = v.z; sg[4 * q + 3][r][c] = v.w;
        }
        __syncthreads();
#pragma unroll 2
        for (int co = 0; co < kCoChunk; co++) {
#pragma unroll
            for (int ky = 0; ky < 3; ky++) {
                // output row ty reads g row ty + 1 - ky (+1 halo); columns tx + p + 1 - kx (+1 halo) -> tx .. tx + 9
                const float *row = &sg[co][ty + 2 - ky][tx];
                const float4 g0 = *reinterpret_cast<const float4 *>(row);
                const float4 g1 = *reinterpret_cast<const float4 *>(row + 4);
                const float2 g2 = *reinterpret_cast<const float2 *>(row + 8);
                const float gv[10] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w, g2.x, g2.y};
#pragma unroll
                for (int kx = 0; kx < 3; kx++) {
                    const float4 wv = *reinterpret_cast<const float4 *>(sw[c0 + co][ky * 3 + kx]);
#pragma unroll
                    for (int p = 0; p < kBwdPx; p++) {
                        const float g = gv[p + 2 - kx];
                        acc[p][0] += g * wv.x; acc[p][1] += g * wv.y; acc[p][2] += g * wv.z;
                    }
                }
            }
        }
    }
    const int y = y0 + ty;
    if (y >= a.H) return;
    float *D = a.dx + (((long long)n * a.H + y) * a.W) * kCin;
#pragma unroll
    for (int p = 0; p < kBwdPx; p++) {
        const int x = x0 + tx + p;
        if (x < a.W) { D[x * 3 + 0] = acc[p][0]; D[x * 3 + 1] = acc[p][1]; D[x * 3 + 2] = acc[p][2]; }
    }
}

}  // namespace

int gom_conv_first_forward_tc(const GomConvFirstArgs *p, cudaStream_t stream);      // conv_first_tc.cu
int gom_conv_first_backward_tc(const GomConvFirstArgs *p, cudaStream_t stream);

extern "C" int gom_conv_first_forward(const GomConvFirstArgs *p, gom_stream_t stream_) {
    GOM_REQUIRE(p != nullptr, "args");
    GOM_REQUIRE(p->n_images > 0 && p->n_images <= 65535 && p->height > 0 && p->width > 0, "sizes");
    GOM_REQUIRE(p->x && p->weight && p->bias && p->out, "null pointer");
    if (p->use_tensor_cores) return gom_conv_first_forward_tc(p, (cudaStream_t)stream_);
    GOM_REQUIRE(((uintptr_t)p->out % 8) == 0, "out must be 8-byte aligned");
    cudaStream_t stream = (cudaStream_t)stream_;
    ConvFwdDev a{p->n_images, p->height, p->width, p->x, p->weight, p->bias, p->out};
    dim3 grid(gom_div_up(a.H, kFwdRows), a.N);
    gom_prof_begin(GOM_PROF_CONV_FIRST_FWD, stream);
    k_conv_first_fwd<<<grid, 256, 0, stream>>>(a);
    GOM_LAUNCH_CHECK();
    gom_prof_end(GOM_PROF_CONV_FIRST_FWD, stream);
    return GOM_OK;
}

extern "C" int gom_conv_first_backward(const GomConvFirstArgs *p, gom_stream_t stream_) {
    GOM_REQUIRE(p != nullptr, "args");
    GOM_REQUIRE(p->n_images > 0 && p->n_images <= 65535 && p->height > 0 && p->width > 0, "sizes");
    GOM_REQUIRE(p->dL_dout && p->weight && p->dL_dx, "null pointer");
    if (p->use_tensor_cores) return gom_conv_first_backward_tc(p, (cudaStream_t)stream_);
    GOM_REQUIRE(p->act == nullptr, "the fused ReLU backward (act) needs use_tensor_cores");
    cudaStream_t stream = (cudaStream_t)stream_;
    ConvBwdDev a{p->n_images, p->height, p->width, p->dL_dout, p->weight, p->dL_dx};
    dim3 grid(gom_div_up(a.W, kBwdCols), gom_div_up(a.H, kBwdRows), a.N);
    GOM_REQUIRE(grid.y <= 65535, "image too tall");
    gom_prof_begin(GOM_PROF_CONV_FIRST_BWD, stream);
    k_conv_first_bwd<<<grid, kBwdThreads, 0, stream>>>(a);
    GOM_LAUNCH_CHECK();
    gom_prof_end(GOM_PROF_CONV_FIRST_BWD, stream);
    return GOM_OK;
}

extern "C" size_t gom_sizeof_conv_first_args(void) { return sizeof(GomConvFirstArgs); }
