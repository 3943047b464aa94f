// narrow_linear.cu — a Linear layer with very few outputs over a tall batch of rows (gom_narrow_linear_*): the last layer of
// the non-rigid deformation MLP (reference models/modules/non_rigid_module.py:112-118: 128 -> 3 offsets per vertex, 120 k rows
// per step).  cuBLAS runs y = x W^T, dx = g W and dW = g^T x of this shape on SIMT sgemm kernels at a few percent of the HBM
// bandwidth the three of them need (0.27 ms per step); here each direction is one pass over x: a warp per row, four columns per
// lane, the (at most 4) weight rows in registers.
//   forward : y[r, j] = sum_c x[r, c] W[j, c] + b[j]
//   backward: dx[r, c] = sum_j g[r, j] W[j, c];  dW[j, c] += sum_r g[r, j] x[r, c];  db[j] += sum_r g[r, j]
#include "gom_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxOut = 4;
constexpr int kMaxIn = 256;                 // two float4 per lane

template <int VEC>                          // float4 per lane: columns 4 (lane + 32 v) .. + 3
__global__ void __launch_bounds__(kThreads) k_narrow_linear_fwd(GomNarrowLinearArgs a) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int C4 = a.c_in / 4;
    float4 w[kMaxOut][VEC];
#pragma unroll
    for (int j = 0; j < kMaxOut; j++)
#pragma unroll
        for (int v = 0; v < VEC; v++) {
            const int c = lane + 32 * v;
            w[j][v] = (j < a.n_out && c < C4) ? __ldg(reinterpret_cast<const float4 *>(a.weight + (long long)j * a.c_in) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    const long long stride = (long long)gridDim.x * (kThreads / 32);
    for (long long r = (long long)blockIdx.x * (kThreads / 32) + warp; r < a.rows; r += stride) {
        const float4 *x = reinterpret_cast<const float4 *>(a.x + r * a.c_in);
        float acc[kMaxOut] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int v = 0; v < VEC; v++) {
            const int c = lane + 32 * v;
            if (c < C4) {
                const float4 xv = __ldg(x + c);
#pragma unroll
                for (int j = 0; j < kMaxOut; j++) acc[j] += xv.x * w[j][v].x + xv.y * w[j][v].y + xv.z * w[j][v].z + xv.w * w[j][v].w;
            }
        }
#pragma unroll
        for (int j = 0; j < kMaxOut; j++) acc[j] = warp_sum(acc[j]);
        if (lane < a.n_out) {
            float y = lane == 0 ? acc[0] : lane == 1 ? acc[1] : lane == 2 ? acc[2] : acc[3];
            if (a.bias) y += a.bias[lane];
            a.y[r * a.n_out + lane] = y;
        }
    }
}

template <int VEC>
__global__ void __launch_bounds__(kThreads) k_narrow_linear_bwd(GomNarrowLinearArgs a) {
    __shared__ float s_w[kThreads / 32][kMaxOut][kMaxIn];
    __shared__ float s_b[kThreads / 32][kMaxOut];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int C4 = a.c_in / 4;
    float4 w[kMaxOut][VEC], gw[kMaxOut][VEC];
    float gb[kMaxOut] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < kMaxOut; j++)
#pragma unroll
        for (int v = 0; v < VEC; v++) {
            const int c = lane + 32 * v;
            w[j][v] = (j < a.n_out && c < C4) ? __ldg(reinterpret_cast<const float4 *>(a.weight + (long long)j * a.c_in) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            gw[j][v] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    const long long stride = (long long)gridDim.x * (kThreads / 32);
    for (long long r = (long long)blockIdx.x * (kThreads / 32) + warp; r < a.rows; r += stride) {
        float g[kMaxOut];
#pragma unroll
        for (int j = 0; j < kMaxOut; j++) g[j] = j < a.n_out ? __ldg(a.g_y + r * a.n_out + j) : 0.f;
#pragma unroll
        for (int j = 0; j < kMaxOut; j++) gb[j] += g[j];
        const float4 *x = reinterpret_cast<const float4 *>(a.x + r * a.c_in);
        float4 *dx = a.g_x ? reinterpret_cast<float4 *>(a.g_x + r * a.c_in) : nullptr;
#pragma unroll
        for (int v = 0; v < VEC; v++) {
            const int c = lane + 32 * v;
            if (c < C4) {
                const float4 xv = __ldg(x + c);
                float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int j = 0; j < kMaxOut; j++) {
                    o.x += g[j] * w[j][v].x; o.y += g[j] * w[j][v].y; o.z += g[j] * w[j][v].z; o.w += g[j] * w[j][v].w;
                    gw[j][v].x += g[j] * xv.x; gw[j][v].y += g[j] * xv.y; gw[j][v].z += g[j] * xv.z; gw[j][v].w += g[j] * xv.w;
                }
                if (dx) dx[c] = o;
            }
        }
    }
    // block reduction of the weight / bias gradient partials, one atomic per element and block
#pragma unroll
    for (int j = 0; j < kMaxOut; j++) {
#pragma unroll
        for (int v = 0; v < VEC; v++) {
            const int c = 4 * (lane + 32 * v);
            if (c < a.c_in) { s_w[warp][j][c] = gw[j][v].x; s_w[warp][j][c + 1] = gw[j][v].y; s_w[warp][j][c + 2] = gw[j][v].z; s_w[warp][j][c + 3] = gw[j][v].w; }
        }
        if (lane == 0) s_b[warp][j] = gb[j];        // every lane of a warp holds the same row sums
    }
    __syncthreads();
    for (int i = threadIdx.x; i < a.n_out * a.c_in; i += kThreads) {
        const int j = i / a.c_in, c = i - j * a.c_in;
        float s = 0.f;
#pragma unroll
        for (int wq = 0; wq < kThreads / 32; wq++) s += s_w[wq][j][c];
        atomicAdd(a.g_weight + i, s);
    }
    if (a.g_bias && threadIdx.x < a.n_out) {
        float s = 0.f;
#pragma unroll
        for (int wq = 0; wq < kThreads / 32; wq++) s += s_b[wq][threadIdx.x];
        atomicAdd(a.g_bias + threadIdx.x, s);
    }
}

int check(const GomNarrowLinearArgs *p) {
    GOM_REQUIRE(p != nullptr, "args");
    GOM_REQUIRE(p->rows > 0, "rows");
    GOM_REQUIRE(p->n_out >= 1 && p->n_out <= kMaxOut, "n_out must be 1 .. 4");
    GOM_REQUIRE(p->c_in >= 4 && p->c_in % 4 == 0 && p->c_in <= kMaxIn, "c_in must be a multiple of 4, at most 256");
    GOM_REQUIRE(p->x && p->weight, "null pointer");
    GOM_REQUIRE(((uintptr_t)p->x % 16) == 0 && ((uintptr_t)p->weight % 16) == 0 && ((uintptr_t)p->g_x % 16) == 0, "16-byte alignment");
    return GOM_OK;
}

int grid_for(long long rows) {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long want = 4ll * sms, need = (rows + kThreads / 32 - 1) / (kThreads / 32);
    return (int)(need < want ? need : want);
}

}  // namespace

extern "C" int gom_narrow_linear_forward(const GomNarrowLinearArgs *p, gom_stream_t stream) {
    if (int rc = check(p)) return rc;
    GOM_REQUIRE(p->y, "null output");
    if (p->c_in <= 128) k_narrow_linear_fwd<1><<<grid_for(p->rows), kThreads, 0, (cudaStream_t)stream>>>(*p);
    else k_narrow_linear_fwd<2><<<grid_for(p->rows), kThreads, 0, (cudaStream_t)stream>>>(*p);
    GOM_LAUNCH_CHECK();
    return GOM_OK;
}

extern "C" int gom_narrow_linear_backward(const GomNarrowLinearArgs *p, gom_stream_t stream_) {
    if (int rc = check(p)) return rc;
    GOM_REQUIRE(p->g_y && p->g_weight, "null gradient pointer");
    cudaStream_t stream = (cudaStream_t)stream_;
    GOM_CUDA(cudaMemsetAsync(p->g_weight, 0, sizeof(float) * (size_t)p->n_out * p->c_in, stream));
    if (p->g_bias) GOM_CUDA(cudaMemsetAsync(p->g_bias, 0, sizeof(float) * (size_t)p->n_out, stream));
    if (p->c_in <= 128) k_narrow_linear_bwd<1><<<grid_for(p->rows), kThreads, 0, stream>>>(*p);
    else k_narrow_linear_bwd<2><<<grid_for(p->rows), kThreads, 0, stream>>>(*p);
    GOM_LAUNCH_CHECK();
    return GOM_OK;
}

extern "C" size_t gom_sizeof_narrow_linear_args(void) { return sizeof(GomNarrowLinearArgs); }
